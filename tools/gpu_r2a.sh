#!/bin/bash
# round-2 GPU pass A: parity tests, first pipeline bench line, host-side profile of the window flow
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/r2a_prof_pipeline.json > gpurun_out/r2a_bench_pipeline.json 2> gpurun_out/r2a_bench_pipeline.err
echo "bench exit $?"; tail -c 3000 gpurun_out/r2a_bench_pipeline.json; tail -5 gpurun_out/r2a_bench_pipeline.err
python tools/pipeline_profile.py 64 64 > gpurun_out/r2a_profile.txt 2>&1
head -60 gpurun_out/r2a_profile.txt
