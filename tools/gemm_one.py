"""One tcgen05 GEMM shape through rdb_debug_gemm (for ncu).  usage: gemm_one.py M N K act res [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
M, N, K, act, res = (int(a) for a in sys.argv[1:6])
os.environ["RDB_DEBUG_GEMM_REPS"] = sys.argv[6] if len(sys.argv) > 6 else "3"
import numpy as np  # noqa: E402
from rapiddoc_b200 import _lib  # noqa: E402

lib = _lib.load()
rng = np.random.default_rng(0)
A = rng.standard_normal((M, K)).astype(np.float32)
W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
b = rng.standard_normal(N).astype(np.float32)
R = rng.standard_normal((M, N)).astype(np.float32) if res else None
out = np.empty((M, N), np.float32)
_lib.check(lib.rdb_debug_gemm(0, 1, 0, A.ctypes.data, W.ctypes.data, b.ctypes.data, R.ctypes.data if res else None, M, N, K, act, out.ctypes.data))
print("ok", float(np.abs(out).mean()))
