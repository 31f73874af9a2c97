#!/bin/bash
# A/B of the shared-address-space smem base (LDS/STS instead of generic LD/ST in the fused kernels), VERDICT r1 item 5a
mkdir -p gpurun_out
python bench.py --workload det --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_det_default.json 2>/dev/null
python bench.py --workload rec --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_rec_default.json 2>/dev/null
RDB_SMEM_BASE=shared python -m rapiddoc_b200.build --force > gpurun_out/r2e_build.log 2>&1; echo "build exit $?"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
python bench.py --workload det --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_det_shared.json 2>/dev/null
python bench.py --workload rec --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_rec_shared.json 2>/dev/null
python - <<'PY'
import json
for w in ("det","rec"):
    for v in ("default","shared"):
        d=json.load(open(f"gpurun_out/r2e_{w}_{v}.json")); print(w, v, round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
PY
