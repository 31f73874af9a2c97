#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_formula.py tests/test_gpu_fused.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -3
for mode in parallel serial; do
  RDB_FORMULA_QKV=$mode python bench.py --workload formula --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_formula_$mode.json 2> gpurun_out/r2t_formula_$mode.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2t_formula_$mode.json")); print("$mode", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "decoder ms", round(d["roofline"]["decoder_ms_per_step"],2), "enc", round(d["roofline"]["encoder_ms_per_step"],2))
PY
done
python bench.py --workload formula --steps 5 --warmup 3 > gpurun_out/r2t_formula.json 2> gpurun_out/r2t_formula.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2t_formula.json")); print(round(d["value"],1), d["cpu_baseline"])
PY
python tools/formula_decode_profile.py 2>&1 | tail -25
