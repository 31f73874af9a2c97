#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_onnx_run.py -m gpu -q --timeout 300 -p no:cacheprovider -k "two_devices or second_device" 2>&1 | tail -4
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pipeline', round(d['value'],1), 'fp32', d['fp32_exact'])" 
