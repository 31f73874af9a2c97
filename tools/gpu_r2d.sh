#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_facade.py tests/test_formula.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
python tools/pipeline_profile.py 64 256 2>&1 | head -16
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2d_bench_stream.json 2> gpurun_out/r2d_bench_stream.err; echo "exit $?"; tail -2 gpurun_out/r2d_bench_stream.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-stream > gpurun_out/r2d_bench_nostream.json 2> gpurun_out/r2d_bench_nostream.err; echo "exit $?"
python - <<'PY'
import json
for f in ("stream","nostream"):
    d=json.load(open(f"gpurun_out/r2d_bench_{f}.json")); print(f, d["value"], d["e2e"]["value"], d["ms_per_step"])
PY
