#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_onnx_run.py tests/test_gpu_tf32.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -6
timeout 300 python bench.py --workload table --steps 10 --warmup 3 --profile-out gpurun_out/r3a_prof_table.json > gpurun_out/r3a_bench_table.json 2> gpurun_out/r3a_bench_table.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r3a_bench_table.json")); r=d["roofline"]
print("table", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2), d["cpu_baseline"]["token_mismatch_vs_oracle"], d["cpu_baseline"]["max_abs_dprob"], "launches", d["gpu_launches"])
p=json.load(open("gpurun_out/r3a_prof_table.json")); print(round(p["total_ms"],2), [(k["kernel"], round(k["total_ms"],3), k["launches"]) for k in p["kernels"][:5]])
PY
