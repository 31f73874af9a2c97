#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tf32.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -12
RDB_TF32=simple timeout 300 python -m pytest tests/test_gpu_tf32.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
for mode in persistent simple; do
RDB_TF32=$mode timeout 300 python bench.py --workload table --precision tf32 --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r3d_prof_table_$mode.json > gpurun_out/r3d_bench_table_$mode.json 2> gpurun_out/r3d_bench_table_$mode.err
python - <<PY
import json
d=json.load(open("gpurun_out/r3d_bench_table_$mode.json")); r=d["roofline"]
print("$mode", round(d["value"],1), "ms", round(d["ms_per_step"],2), "gemm GB/s", round(r["achieved"]), round(r["frac"],3))
PY
done
