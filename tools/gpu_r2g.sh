#!/bin/bash
# round-2 GPU pass G: ONNX executor tests (orientation classifier / seal detector), clock-sampler A/B on the pipeline bench
mkdir -p gpurun_out
python -m pytest tests/test_onnx_run.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/r2g_pytest.log
for mode in nvml smi nvml smi; do
  RDB_BENCH_CLOCKS=$mode python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2g_bench_$mode.json 2> gpurun_out/r2g_bench_$mode.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench_$mode.json")); print("$mode", round(d["value"],1), round(d["e2e"]["value"],1), d["clocks"])
PY
done
