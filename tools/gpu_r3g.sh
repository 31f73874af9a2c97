#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3g_pipe_n2.json 2> gpurun_out/r3g_pipe_n2.err
echo "pipeline N=2 exit $?; stdout lines: $(wc -l < gpurun_out/r3g_pipe_n2.json)"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r3g_pipe_n2.json")); print(round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["traffic"], d["fp32_exact"]["value"] if d.get("fp32_exact") else None, d["n_gpus"])
PY
