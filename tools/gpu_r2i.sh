#!/bin/bash
# round-2 GPU pass I: fused conv epilogues in the ONNX executor + cluster/DSMEM SLA decode — tests, A/B against the one-CTA kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_onnx_run.py tests/test_table_match.py tests/test_callers.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/r2i_pytest.log
RDB_SLA=cta timeout 600 python -m pytest tests/test_onnx_run.py -m gpu -q --timeout 300 -p no:cacheprovider -k slanet 2>&1 | tail -3
for mode in cluster cta; do
RDB_SLA=$mode timeout 300 python bench.py --workload table --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2i_prof_table_$mode.json > gpurun_out/r2i_bench_table_$mode.json 2> gpurun_out/r2i_bench_table_$mode.err
echo "table bench $mode exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2i_bench_table_$mode.json")); print("$mode", round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "backbone_ms", round(d["roofline"]["backbone_ms"],2), "steps", d["roofline"]["decode_steps"])
PY
done
timeout 300 python bench.py --workload table --steps 5 --warmup 3 > gpurun_out/r2i_bench_table.json 2> gpurun_out/r2i_bench_table.err; head -c 600 gpurun_out/r2i_bench_table.json; echo
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i_bench_table.json")); print(d["cpu_baseline"])
PY
