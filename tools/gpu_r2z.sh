#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tf32.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -25
for pr in fp32 tf32; do
timeout 300 python bench.py --workload table --precision $pr --steps 10 --warmup 3 --profile-out gpurun_out/r2z_prof_table_$pr.json > gpurun_out/r2z_bench_table_$pr.json 2> gpurun_out/r2z_bench_table_$pr.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2z_bench_table_$pr.json")); r=d["roofline"]
print("$pr", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2), d["cpu_baseline"]["token_mismatch_vs_oracle"], d["cpu_baseline"]["max_abs_dprob"], d["cpu_baseline"]["max_abs_dbox"], "gemm GB/s", round(r["achieved"]), round(r["frac"],3))
PY
done
