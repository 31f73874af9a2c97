#!/bin/bash
# copy the judged artefacts of the last tools/collect_gpu.sh run from gpurun_out/ (scratch) into profiles/ (tracked)
set -e
cd "$(dirname "$0")/.."
R=${1:-r01}
mkdir -p profiles
for f in bench_det_fp16 bench_rec_fp16 bench_det_fp32 bench_ref prof_det_fp16 prof_rec_fp16; do
  [ -s gpurun_out/$f.json ] && cp gpurun_out/$f.json profiles/${R}_$f.json
done
for f in launches_det launches_rec; do
  [ -s gpurun_out/$f.csv ] && grep -v "^==" gpurun_out/$f.csv > profiles/${R}_ncu_$f.csv
done
[ -s gpurun_out/pytest_gpu.log ] && cp gpurun_out/pytest_gpu.log profiles/${R}_pytest_gpu.log
reps=""
for f in ncu_det_top ncu_rec_top; do
  [ -s gpurun_out/$f.ncu-rep ] && reps="$reps gpurun_out/$f.ncu-rep"
done
[ -n "$reps" ] && python tools/summarize_ncu.py profiles/${R}_ncu_summary $reps > /dev/null
python tools/ncu_traffic.py profiles/${R}_ncu_summary.json > profiles/ncu_traffic.json
ls -la profiles
