"""Epilogue cost of the tcgen05 GEMM: same shape with act none / relu / gelu / silu (rdb_debug_gemm + rdb_profile_*).
usage: RDB_DEBUG_GEMM_REPS=20 python tools/gemm_act_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RDB_DEBUG_GEMM_REPS", "20")
import numpy as np  # noqa: E402
from rapiddoc_b200 import _lib  # noqa: E402

lib = _lib.load()
rng = np.random.default_rng(0)
ACTS = {"none": 0, "relu": 1, "gelu": 2, "silu": 3}
for (M, N, K, res) in [(524288, 96, 48, 0), (524288, 48, 96, 1), (131072, 192, 96, 0), (32768, 384, 192, 0)]:
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    R = rng.standard_normal((M, N)).astype(np.float32) if res else None
    out = np.empty((M, N), np.float32)
    for name, act in ACTS.items():
        _lib.profile(True); _lib.profile_reset()
        _lib.check(lib.rdb_debug_gemm(0, 1, 0, A.ctypes.data, W.ctypes.data, b.ctypes.data, R.ctypes.data if res else None, M, N, K, act,
                                      out.ctypes.data))
        prof = _lib.profile_dump(); _lib.profile(False)
        for k, v in prof.items():
            byts = M * (K + N + (N if res else 0)) * 2
            us = v[0] / v[1] * 1e3
            print(f"M={M} K={K} N={N} res={res} act={name:5s} {k:40s} n={v[1]} avg_us={us:7.1f}  {byts / us / 1e3:7.1f} GB/s")
