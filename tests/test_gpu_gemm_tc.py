"""GPU unit tests of the tcgen05/TMA GEMM (gemm_tc.cuh) through the diagnostic C-ABI entry,
against a float64 numpy reference on the same fp16-rounded inputs, for every (K, N) shape the
two networks use plus ragged edges."""
import numpy as np
import pytest

from rapiddoc_b200 import _lib

pytestmark = pytest.mark.gpu
ACTS = {0: lambda x: x, 1: lambda x: np.maximum(x, 0), 3: lambda x: x / (1 + np.exp(-x))}


def gelu(x):
    from math import erf
    return 0.5 * x * (1 + np.vectorize(erf)(x / np.sqrt(2)))


def run(M, N, K, act=0, bias=True, res=False, use_tc=1, seed=0, mode=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32) if bias else None
    R = rng.standard_normal((M, N)).astype(np.float32) if res else None
    out = np.empty((M, N) if mode == 0 else (M, 2), np.float32)
    lib = _lib.load()
    _lib.check(lib.rdb_debug_gemm(0, use_tc, mode, A.ctypes.data, W.ctypes.data, b.ctypes.data if bias else None,
                                  R.ctypes.data if res else None, M, N, K, act, out.ctypes.data))
    Ah, Wh = A.astype(np.float16).astype(np.float64), W.astype(np.float16).astype(np.float64)
    ref = Ah @ Wh.T + (b.astype(np.float64) if bias else 0)
    return out, ref, (R.astype(np.float16).astype(np.float64) if res else None)


SHAPES = [(128, 48, 24), (300, 96, 48), (1000, 48, 96), (257, 192, 96), (129, 96, 192), (640, 384, 192), (513, 192, 384),
          (200, 768, 384), (131, 384, 768), (77, 24, 96), (4096, 96, 48), (333, 120, 384), (90, 360, 120), (90, 240, 120),
          (90, 120, 240), (1, 96, 48), (128 * 150 + 5, 96, 96)]


@pytest.mark.parametrize("use_tc", [1, 0])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_shapes(use_tc, M, N, K):
    out, ref, _ = run(M, N, K, use_tc=use_tc, seed=M + N + K)
    err = np.abs(out - ref).max()
    assert err <= 2e-3 * max(1.0, np.abs(ref).max()), err     # fp16 output rounding only


@pytest.mark.parametrize("act", [1, 2, 3])
def test_gemm_epilogue_act_res(act):
    out, ref, R = run(700, 192, 96, act=act, res=True, seed=act)
    want = (gelu(ref) if act == 2 else ACTS[act](ref)) + R
    assert np.abs(out - want).max() <= 4e-3 * max(1.0, np.abs(want).max())


def test_gemm_no_bias():
    out, ref, _ = run(260, 24, 96, bias=False)
    assert np.abs(out - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("use_tc", [1, 0])
def test_ctc_epilogue(use_tc):
    M, N, K = 600, 18710, 120
    out, ref, _ = run(M, N, K, use_tc=use_tc, mode=1, seed=5)
    ids = out[:, 0].astype(np.int64)
    srt = np.sort(ref, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    bad = (ids != ref.argmax(1)) & (margin > 1e-3)
    assert not bad.any()
    p = 1.0 / np.exp(ref - ref.max(1, keepdims=True)).sum(1)
    assert np.abs(out[:, 1] - p).max() <= 2e-3
