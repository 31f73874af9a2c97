#!/bin/bash
# round-2 GPU pass H: ONNX executor (orientation / seal / SLANet backbone) + SLA decode kernel tests
mkdir -p gpurun_out
python -m pytest tests/test_onnx_run.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2h_pytest.log 2>&1
echo "pytest exit $?"; tail -40 gpurun_out/r2h_pytest.log
python - <<'PY' 2>&1 | tail -20
import sys, time; sys.path.insert(0, ".")
import numpy as np, torch
from rapiddoc_b200 import table, synth
imgs = [synth.table_image(i, 3 + i % 6, 2 + i % 4, 300 + 10 * (i % 5), 400 + 16 * (i % 7), lines=(i % 3 != 0)) for i in range(32)]
ts = table.B200TableStructurer(device=0)
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    structs, cells = ts(imgs)
    torch.cuda.synchronize(); dt = time.time() - t
    print("structurer b32: %.1f ms, steps %d, launches %d" % (dt * 1e3, ts.session.last_steps, ts.session.launches))
x, shapes = ts.preprocess_op(imgs); x = np.asarray(x)
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    ts.session(x)
    torch.cuda.synchronize(); print("session b32: %.1f ms" % ((time.time() - t) * 1e3))
from rapiddoc_b200 import _lib
lib = _lib.load()
lib.rdb_profile_enable(1)
ts.session(x)
torch.cuda.synchronize()
import ctypes, json
n = lib.rdb_profile_dump(None, 0)
buf = ctypes.create_string_buffer(n + 16); lib.rdb_profile_dump(buf, n + 16)
prof = json.loads(buf.value.decode())
tot = sum(v[0] for v in prof.values())
print("profiled ops total %.2f ms" % tot)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:12]: print("  %-50s %8.3f ms x%d" % (k, v[0], v[1]))
PY
python bench.py --workload table --steps 5 --warmup 3 --profile-out gpurun_out/r2h_prof_table.json > gpurun_out/r2h_bench_table.json 2> gpurun_out/r2h_bench_table.err
echo "table bench exit $?"; head -c 2500 gpurun_out/r2h_bench_table.json; echo; tail -3 gpurun_out/r2h_bench_table.err
