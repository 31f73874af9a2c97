#!/bin/bash
# round-2 GPU pass O: full validation of HEAD — whole gpu test suite, smoke, default bench (+profile), formula / table benches, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r3b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3b_pytest.log; tail -4 gpurun_out/r3b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3b_smoke.log 2>&1; tail -2 gpurun_out/r3b_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3b_bench_reference.json 2> gpurun_out/r3b_bench_reference.err
echo "reference exit $?"; head -c 300 gpurun_out/r3b_bench_reference.json; echo
python bench.py --profile-out gpurun_out/r3b_prof_pipeline.json > gpurun_out/r3b_bench_pipeline.json 2> gpurun_out/r3b_bench_pipeline.err
echo "bench exit $?"; head -c 400 gpurun_out/r3b_bench_pipeline.json; echo; tail -3 gpurun_out/r3b_bench_pipeline.err
python bench.py --workload formula --steps 5 --warmup 3 --profile-out gpurun_out/r3b_prof_formula_enc.json > gpurun_out/r3b_bench_formula_fp16.json 2> gpurun_out/r3b_bench_formula_fp16.err
echo "formula fp16 exit $?"; head -c 300 gpurun_out/r3b_bench_formula_fp16.json; echo
python bench.py --workload table --steps 10 --warmup 3 --profile-out gpurun_out/r3b_prof_table.json > gpurun_out/r3b_bench_table.json 2> gpurun_out/r3b_bench_table.err
echo "table exit $?"; head -c 300 gpurun_out/r3b_bench_table.json; echo
python bench.py --workload det --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r3b_bench_det.json 2> gpurun_out/r3b_bench_det.err; echo "det exit $?"
python bench.py --workload rec --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r3b_bench_rec.json 2> gpurun_out/r3b_bench_rec.err; echo "rec exit $?"
python - <<'PY'
import json
for n in ("pipeline","formula_fp16","table","det","rec"):
    try:
        d=json.load(open(f"gpurun_out/r3b_bench_{n}.json")); print(n, round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "roofline", d["roofline"].get("kernel"), d["roofline"].get("frac"))
    except Exception as e: print(n, "failed", e)
PY
