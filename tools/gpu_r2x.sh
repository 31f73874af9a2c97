#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_gpu_pipeline.py tests/test_onnx_run.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2x_prof_pipeline.json > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2x_bench.json")); print("pipeline", round(d["value"],1), round(d["e2e"]["value"],1), d["secondary"])
p=json.load(open("gpurun_out/r2x_prof_pipeline.json"))
fam={}
for k in p["kernels"]:
    f=k["kernel"].split("[")[0]; fam.setdefault(f,[0,0]); fam[f][0]+=k["total_ms"]; fam[f][1]+=k["launches"]
print("total", round(p["total_ms"],2), {k:(round(v[0],2),v[1]) for k,v in sorted(fam.items(), key=lambda kv:-kv[1][0])[:4]})
for k in p["kernels"]:
    if k["kernel"].startswith("dwconv3x3_tiled") and k["total_ms"]>0.25: print(k["kernel"], round(k["total_ms"],3), k["launches"])
PY
