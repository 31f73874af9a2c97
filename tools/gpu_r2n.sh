#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
from rapiddoc_b200.parallel import gpu_ideal_cores, plan_rank_cores
import os
ideal = gpu_ideal_cores(4)
print("ideal", [sorted(s)[:4] + ["..."] + sorted(s)[-2:] + [len(s)] for s in ideal] if ideal else None)
print("plan", plan_rank_cores(sorted(os.sched_getaffinity(0)), 4, ideal))
PY
nvidia-smi topo -m 2>/dev/null | head -12
for n in 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2n_bench_n$n.json 2> gpurun_out/r2n_bench_n$n.err
  echo "N=$n exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2n_bench_n$n.json")); print("N=$n", round(d["value"],1), round(d["e2e"]["value"],1), d["config"].get("host_cores_per_rank"))
PY
done
