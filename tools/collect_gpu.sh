#!/bin/bash
# run ON the GPU box (under gpurun): every artefact that tools/collect_profiles.sh later copies into profiles/
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/prof_det_fp16.json > gpurun_out/bench_det_fp16.json 2> gpurun_out/bench_det_fp16.err
python bench.py --workload rec --steps 20 --warmup 3 --profile-out gpurun_out/prof_rec_fp16.json > gpurun_out/bench_rec_fp16.json 2> gpurun_out/bench_rec_fp16.err
python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_det_fp32.json 2> gpurun_out/bench_det_fp32.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_det.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_det.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_rec.csv python bench.py --workload rec --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"stem_planar|head_planar|mlp_tc_kernel<48, 48|mlp_big|dwconv_tiled_h2_kernel<7|stem1_tc" -c 9 -o gpurun_out/ncu_det_top python tools/run_once.py 32 > gpurun_out/ncu_det_top.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mlp_big" -c 1 -o gpurun_out/ncu_rec_top python tools/rec_once.py > gpurun_out/ncu_rec_top.log 2>&1
cat gpurun_out/pytest_gpu.log; cut -c1-330 gpurun_out/bench_det_fp16.json; cut -c1-250 gpurun_out/bench_rec_fp16.json; cut -c1-200 gpurun_out/bench_det_fp32.json; cut -c1-200 gpurun_out/bench_ref.json
