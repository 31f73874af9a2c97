#!/bin/bash
# round-2 GPU pass C: full gpu test suite, default bench line, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c_pytest.log; tail -5 gpurun_out/r2c_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; tail -2 gpurun_out/r2c_smoke.log
python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r2c_prof_pipeline.json > gpurun_out/r2c_bench_pipeline.json 2> gpurun_out/r2c_bench_pipeline.err
echo "bench exit $?"; head -c 1500 gpurun_out/r2c_bench_pipeline.json; echo; tail -3 gpurun_out/r2c_bench_pipeline.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_ncu_launches_pipeline.csv python tools/ncu_pipeline_step.py 32 > gpurun_out/r2c_ncu_launches.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/r2c_ncu_launches_pipeline.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"mlp_big|mlp_tc|gemm_tc_kernel|dwconv_tiled" -c 24 -o gpurun_out/r2c_ncu_top python tools/ncu_pipeline_step.py 32 > gpurun_out/r2c_ncu_top.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/r2c_ncu_top.ncu-rep
