#!/bin/bash
# final check of HEAD: whole gpu suite, smoke, default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r3e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3e_pytest.log; tail -4 gpurun_out/r3e_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3e_smoke.log 2>&1; tail -2 gpurun_out/r3e_smoke.log
python bench.py > gpurun_out/r3e_bench_pipeline.json 2> gpurun_out/r3e_bench_pipeline.err
echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r3e_bench_pipeline.json")); print(round(d["value"],1), round(d["e2e"]["value"],1), d["fp32_exact"]["value"], d["fp32_exact"]["parity"]["text_mismatch"], d["parity"]["text_mismatch"], d["clocks"])
PY
