#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_onnx_run.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload table --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_table_n2.json 2> gpurun_out/r2q_table_n2.err
echo "table N=2 exit $?; stdout lines: $(wc -l < gpurun_out/r2q_table_n2.json)"; head -c 200 gpurun_out/r2q_table_n2.json; echo; grep -c "NCCL version" gpurun_out/r2q_table_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2q_pipe_n2.json 2> gpurun_out/r2q_pipe_n2.err
echo "pipeline N=2 exit $?; stdout lines: $(wc -l < gpurun_out/r2q_pipe_n2.json)"; head -c 160 gpurun_out/r2q_pipe_n2.json; echo
