#!/bin/bash
# round-2 GPU pass F: full gpu test suite, smoke, default bench line (+profile), formula bench fp16/fp32, det/rec secondaries
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2f_pytest.log; tail -5 gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r2f_prof_pipeline.json > gpurun_out/r2f_bench_pipeline.json 2> gpurun_out/r2f_bench_pipeline.err
echo "bench exit $?"; head -c 1200 gpurun_out/r2f_bench_pipeline.json; echo; tail -3 gpurun_out/r2f_bench_pipeline.err
python bench.py --workload formula --steps 5 --warmup 3 > gpurun_out/r2f_bench_formula_fp16.json 2> gpurun_out/r2f_bench_formula_fp16.err
echo "formula fp16 exit $?"; head -c 900 gpurun_out/r2f_bench_formula_fp16.json; echo
python bench.py --workload formula --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_formula_fp32.json 2> gpurun_out/r2f_bench_formula_fp32.err
echo "formula fp32 exit $?"; head -c 600 gpurun_out/r2f_bench_formula_fp32.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
echo "reference exit $?"; head -c 600 gpurun_out/r2f_bench_reference.json; echo
