#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import os
from rapiddoc_b200.parallel import smt_order, plan_rank_cores
o = smt_order(os.sched_getaffinity(0)); print("smt order", o)
print("plan4", plan_rank_cores(o, 4))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2p_bench_n4.json 2> gpurun_out/r2p_bench_n4.err
echo "N=4 exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2p_bench_n4.json")); print("N=4", round(d["value"],1), round(d["e2e"]["value"],1), d["config"].get("host_cores_per_rank"), d["ms_per_step"])
PY
CUDA_VISIBLE_DEVICES=0 python bench.py --workload table --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench_table.json 2> gpurun_out/r2p_bench_table.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2p_bench_table.json")); r=d["roofline"]; print("table", round(d["value"],1), {k:r[k] for k in ("kernel","achieved","frac","launches_profiled","avg_launch_us","share_of_step")})
PY
