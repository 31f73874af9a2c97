"""RDB_PREC_TF32: the tcgen05 kind::tf32 GEMM on fp32 storage (csrc/gemm_tf32.cuh) against a float64 product of the inputs with
their mantissas cut to TF32, and the SLANet path in that mode against the fp32 oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rapiddoc_b200 import _lib      # noqa: E402

pytestmark = pytest.mark.gpu


def _tf32(a):
    """fp32 -> the value the tensor core sees: low 13 mantissa bits dropped (tried as truncation and as round-to-nearest)."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32)
    trunc = (u & np.uint32(0xFFFFE000)).view(np.float32)
    rne = ((u + np.uint32(0x0FFF) + ((u >> np.uint32(13)) & np.uint32(1))) & np.uint32(0xFFFFE000)).view(np.float32)
    return trunc, rne


@pytest.mark.parametrize("M,K,N,act,ldc,c_off", [(128, 32, 16, 0, 16, 0), (1000, 36, 48, 6, 48, 0), (4097, 192, 48, 6, 192, 96), (300, 256, 256, 1, 256, 0),
                                                (257, 100, 30, 0, 32, 0), (513, 64, 300, 0, 304, 4), (130, 8, 200, 6, 200, 0)])
def test_gemm_tf32_matches_a_tf32_product(M, K, N, act, ldc, c_off):
    import torch
    rng = np.random.default_rng(M + K + N)
    lda = K + 4
    A = rng.standard_normal((M, lda)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    dA, dW, db = torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda(), torch.from_numpy(bias).cuda()
    out = torch.full((M, ldc), -7.0, dtype=torch.float32, device="cuda")
    lib = _lib.load()
    _lib.check_op(lib.rdb_op_gemm(0, _lib.PREC_TF32, dA.data_ptr(), lda, M, K, dW.data_ptr(), N, db.data_ptr(), act, None, 0, out.data_ptr(), ldc, c_off,
                                  torch.cuda.current_stream().cuda_stream or None, None, 0))
    torch.cuda.synchronize()
    got = out.cpu().numpy()

    def ref(a, w):
        y = a.astype(np.float64) @ w.astype(np.float64).T + bias
        if act == 1:
            y = np.maximum(y, 0)
        elif act == 6:
            y = y * np.clip(y / 6 + 0.5, 0, 1)
        return y
    (at, ar), (wt, wr) = _tf32(A[:, :K]), _tf32(W)
    e_trunc = np.abs(got[:, c_off:c_off + N] - ref(at, wt)).max()
    e_rne = np.abs(got[:, c_off:c_off + N] - ref(ar, wr)).max()
    e_fp32 = np.abs(got[:, c_off:c_off + N] - ref(A[:, :K], W)).max()
    assert min(e_trunc, e_rne) < 2e-5 * np.sqrt(K) + 1e-5, (e_trunc, e_rne, e_fp32)       # fp32 accumulation of exact tf32 products
    assert e_fp32 < 2e-2, e_fp32                                                          # vs plain fp32: tf32 input rounding
    untouched = np.ones(ldc, bool)
    untouched[c_off:c_off + N] = False
    assert np.all(got[:, untouched] == -7.0)                                              # only the slice is written


def test_slanet_in_tf32_mode_keeps_the_tokens():
    from oracle import make_golden_onnx as MG
    from oracle import onnx_ref
    from rapiddoc_b200.table import SlaNetSession
    path = os.path.join(ROOT, "weights", "slanet-1m.onnx")
    _, x, _ = MG.table_inputs()
    s = SlaNetSession(path, 0, precision="tf32")
    loc, probs = s(x)
    rloc, rprobs = onnx_ref.run(path, x)
    assert probs.shape == rprobs.shape and np.array_equal(probs.argmax(-1), rprobs.argmax(-1))
    assert np.abs(probs - rprobs).max() < 3e-2 and np.abs(loc - rloc).max() < 1e-2, (np.abs(probs - rprobs).max(), np.abs(loc - rloc).max())
