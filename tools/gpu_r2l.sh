#!/bin/bash
# round-2 GPU pass L (4 GPUs): torchrun pipeline bench at N=2 and N=4 (rank pinning, streaming warm-up), table workload at N=4
mkdir -p gpurun_out
nproc; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"
for n in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2l_bench_n$n.json 2> gpurun_out/r2l_bench_n$n.err
  echo "N=$n exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2l_bench_n$n.json")); print("N=$n", round(d["value"],1), round(d["e2e"]["value"],1), d["config"].get("host_cores_per_rank"), d["clocks"])
except Exception as e: print("parse failed", e)
PY
  tail -3 gpurun_out/r2l_bench_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --workload table --steps 5 --warmup 3 > gpurun_out/r2l_bench_table_n4.json 2> gpurun_out/r2l_bench_table_n4.err
echo "table N=4 exit $?"; head -c 300 gpurun_out/r2l_bench_table_n4.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 4 --steps 1 --warmup 0 > gpurun_out/r2l_bench_ref_n4.json 2> gpurun_out/r2l_bench_ref_n4.err
echo "reference N=4 exit $?"; head -c 300 gpurun_out/r2l_bench_ref_n4.json; echo
