#!/bin/bash
# round-2 GPU pass J: ncu evidence for the table path and the recogniser kernels; rec-batch A/B on the pipeline bench
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2j_ncu_launches_table.csv python tools/ncu_table_step.py 32 > gpurun_out/r2j_ncu_launches_table.log 2>&1
echo "ncu table launches exit $?"; wc -l gpurun_out/r2j_ncu_launches_table.csv
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sla_decode" -c 1 -o gpurun_out/r2j_ncu_sla python tools/ncu_table_step.py 32 > gpurun_out/r2j_ncu_sla.log 2>&1
echo "ncu sla exit $?"; ls -la gpurun_out/r2j_ncu_sla.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2j_ncu_launches_rec.csv python tools/ncu_rec_step.py > gpurun_out/r2j_ncu_launches_rec.log 2>&1
echo "ncu rec launches exit $?"; wc -l gpurun_out/r2j_ncu_launches_rec.csv
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"mlp_tc|mlp_big|gemm_tc_kernel|dwconv_tiled|se_scale" -c 40 -o /tmp/r2j_ncu_rec python tools/ncu_rec_step.py > gpurun_out/r2j_ncu_rec.log 2>&1
echo "ncu rec full exit $?"; ls -la /tmp/r2j_ncu_rec.ncu-rep
python tools/summarize_ncu.py gpurun_out/r2j_ncu_rec_summary /tmp/r2j_ncu_rec.ncu-rep gpurun_out/r2j_ncu_sla.ncu-rep > gpurun_out/r2j_summarize.log 2>&1; ls -la gpurun_out/
ncu -i /tmp/r2j_ncu_rec.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("/tmp/raw.csv")))
hdr=rows[0]
want=[h for h in hdr if any(k in h for k in ("Kernel Name","gpu__time_duration.sum","smsp__average_warp_latency_issue_stalled","smsp__average_warps_issue_stalled","stall","dram__bytes","l1tex__t_sector_hit_rate","lts__t_sector_hit_rate","achieved_occupancy","sm__warps_active"))]
idx=[hdr.index(h) for h in want]
with open("gpurun_out/r2j_ncu_rec_stalls.csv","w") as f:
    w=csv.writer(f); w.writerow(want); w.writerow([rows[1][i] for i in idx])
    for r in rows[2:]: w.writerow([r[i] for i in idx])
PY
for rb in 256 512; do
  python bench.py --steps 10 --warmup 3 --no-secondary --rec-batch $rb > gpurun_out/r2j_bench_rb$rb.json 2> gpurun_out/r2j_bench_rb$rb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2j_bench_rb$rb.json")); print("rec-batch $rb", round(d["value"],1), round(d["e2e"]["value"],1), d["gpu_launches"], d["parity"])
PY
done
du -sh gpurun_out
