"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI,
against the CPU oracle on the same seeded inputs and against the committed golden
fixtures generated from the reference's own networks.

Tolerances (north_star: bit-exact decisions, stated fp tolerance for raw maps/logits):
  fp32 mode : |prob - oracle| <= 2e-5; bitmap / argmax identical except where the oracle
              itself is within 1e-4 of the decision boundary (counted, must be ~0)
  fp16 mode : |prob - oracle| <= 3e-2; decisions identical except within 2e-2 / logit
              margin 0.25 of the boundary; decoded TEXT must be identical on real lines.
"""
import os

import numpy as np
import pytest

from oracle import nets, ocr_post as P
from rapiddoc_b200 import PREC_FP16, PREC_FP32
from rapiddoc_b200.engine import DetEngine, RecEngine

pytestmark = pytest.mark.gpu

TOL = {PREC_FP32: dict(prob=2e-5, near=1e-4, margin=1e-3, p=1e-4), PREC_FP16: dict(prob=3e-2, near=2e-2, margin=0.25, p=2e-2)}
_ENG = {}


def det_engine(prec):
    if ("det", prec) not in _ENG:
        _ENG[("det", prec)] = DetEngine(device=0, precision=prec)
    return _ENG[("det", prec)]


def rec_engine(prec):
    if ("rec", prec) not in _ENG:
        _ENG[("rec", prec)] = RecEngine(device=0, precision=prec)
    return _ENG[("rec", prec)]


def _check_det(prob, want, prec, bitmap=None, dilate=True):
    t = TOL[prec]
    d = np.abs(prob - want)
    assert d.max() <= t["prob"], f"max |prob diff| {d.max()}"
    if bitmap is not None:
        ref = P.db_bitmap(want, 0.3, dilate)
        seg_flip = (prob > 0.3) != (want > 0.3)
        assert not (seg_flip & (np.abs(want - 0.3) > t["near"])).any(), "bitmap flips away from the threshold"
        if prec == PREC_FP32:
            assert seg_flip.sum() == 0
        own = P.db_bitmap(prob, 0.3, dilate)          # binarise+dilate of OUR prob map on CPU
        assert np.array_equal(bitmap, own), "GPU binarise/dilate != cv2 on the same prob map"
        assert (bitmap != ref).mean() <= (0.0 if prec == PREC_FP32 else 2e-3)   # noise images sit near the threshold


@pytest.mark.parametrize("prec", [PREC_FP32, PREC_FP16])
def test_det_f32_seam_vs_golden(prec, golden_dir):
    g = np.load(os.path.join(golden_dir, "det_randn_2x64x96.npz"))
    prob = det_engine(prec).infer_f32(g["x"])
    assert prob.shape == g["prob"].shape
    _check_det(prob, g["prob"], prec)
    g = np.load(os.path.join(golden_dir, "det_real_192x256.npz"))
    prob = det_engine(prec).infer_f32(g["x"])
    _check_det(prob, g["prob"], prec)


@pytest.mark.parametrize("prec", [PREC_FP32, PREC_FP16])
def test_det_u8_seam_fused_normalise_and_bitmap(prec, golden_dir):
    g = np.load(os.path.join(golden_dir, "det_real_192x256.npz"))
    page = g["page_bgr"]
    for dil in (True, False):
        prob, bm = det_engine(prec).infer_u8(page[None], thresh=0.3, use_dilation=dil)
        _check_det(prob[0], g["prob"][0, 0], prec, bm[0], dil)


@pytest.mark.parametrize("prec", [PREC_FP32, PREC_FP16])
@pytest.mark.parametrize("shape", [(1, 32, 32), (3, 96, 160), (2, 224, 128), (1, 640, 480)])
def test_det_vs_oracle_seeded(prec, shape):
    n, h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    pages = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    pages[:, h // 4: h // 2, w // 8: w // 2] = 255      # flat regions + edges
    pages[:, h // 2:, :] = (pages[:, h // 2:, :] // 64) * 64
    x = np.concatenate([P.det_preprocess(p, limit_side_len=4096) for p in pages])
    want = nets.det_forward(x)[:, 0]
    prob, bm = det_engine(prec).infer_u8(pages)
    for i in range(n):
        _check_det(prob[i], want[i], prec, bm[i], True)
    prob2 = det_engine(prec).infer_f32(x)[:, 0]
    assert np.abs(prob2 - prob).max() <= (1e-6 if prec == PREC_FP32 else 2e-2)


def test_det_chunking_is_invisible():
    rng = np.random.default_rng(0)
    pages = rng.integers(0, 256, (5, 64, 96, 3), dtype=np.uint8)
    e = det_engine(PREC_FP32)
    a, ba = e.infer_u8(pages)
    e.set_chunk_pixels(64 * 96 * 2)
    b, bb = e.infer_u8(pages)
    e.set_chunk_pixels(8 * 1024 * 1024)
    assert np.array_equal(a, b) and np.array_equal(ba, bb)


def test_db_bitmap_entry(golden_dir):
    import ctypes as C
    from rapiddoc_b200 import _lib
    g = np.load(os.path.join(golden_dir, "det_page_img5.npz"))
    prob = np.ascontiguousarray(g["prob"][0, 0])
    for dil in (1, 0):
        out = np.empty(prob.shape, np.uint8)
        _lib.check(_lib.load().rdb_db_bitmap(0, prob.ctypes.data, 1, prob.shape[0], prob.shape[1], 0.3, dil, out.ctypes.data, None))
        assert np.array_equal(out, P.db_bitmap(prob, 0.3, bool(dil)))


def _check_rec(out, logits, prec, T):
    t = TOL[prec]
    ids_w = logits.argmax(2)
    srt = np.sort(logits, axis=2)
    margin = srt[:, :, -1] - srt[:, :, -2]
    mism = out["ids"] != ids_w
    assert not (mism & (margin > t["margin"])).any(), f"argmax flips with margin {margin[mism].max() if mism.any() else 0}"
    if prec == PREC_FP32:
        assert mism.sum() == 0
    lse = np.log(np.exp(logits - logits.max(2, keepdims=True)).sum(2)) + logits.max(2)
    pmax = np.exp(logits.max(2) - lse)
    ok = ~mism
    assert np.abs(out["probs"][ok] - pmax[ok]).max() <= t["p"]
    # collapse + confidence of OUR ids/probs must equal CTCLabelDecode on them
    for b in range(len(ids_w)):
        sel = np.ones(T, bool)
        sel[1:] = out["ids"][b][1:] != out["ids"][b][:-1]
        sel &= out["ids"][b] != 0
        n = int(sel.sum())
        assert out["text_len"][b] == n
        assert np.array_equal(out["text_ids"][b][:n], out["ids"][b][sel])
        assert (out["text_ids"][b][n:] == -1).all()
        want_conf = float(np.mean(out["probs"][b][sel])) if n else 0.0
        assert abs(out["conf"][b] - want_conf) <= 1e-5


@pytest.mark.parametrize("prec", [PREC_FP32, PREC_FP16])
def test_rec_vs_golden_real_lines(prec, golden_dir):
    g = np.load(os.path.join(golden_dir, "rec_real_6lines.npz"))
    e = rec_engine(prec)
    out = e.infer_f32(g["x"])
    T = g["ids"].shape[1]
    assert out["ids"].shape == (6, T)
    mism = out["ids"] != g["ids"]
    assert not (mism & (g["margin"] > TOL[prec]["margin"])).any()
    chars = nets.load_characters()
    texts = ["".join(chars[i] for i in out["text_ids"][b][: out["text_len"][b]]) for b in range(6)]
    assert texts == list(g["texts"])
    assert np.abs(out["conf"] - g["conf"]).max() <= (1e-4 if prec == PREC_FP32 else 1e-2)
    # facade seam: uint8 crops, fused resize_norm_img normalisation + right zero-pad
    crops = [g[f"crop{i}"] for i in range(6)]
    import cv2, math
    W = g["x"].shape[3]
    buf = np.zeros((6, 48, W, 3), np.uint8)
    vw = np.zeros(6, np.int32)
    for i, c in enumerate(crops):
        rw = min(W, int(math.ceil(48 * c.shape[1] / c.shape[0])))
        buf[i, :, :rw] = cv2.resize(c, (rw, 48))
        vw[i] = rw
    out2 = e.infer_u8(buf, vw)
    if prec == PREC_FP32:
        assert np.array_equal(out2["ids"], out["ids"])
        assert np.abs(out2["probs"] - out["probs"]).max() <= 1e-5
    texts2 = ["".join(chars[i] for i in out2["text_ids"][b][: out2["text_len"][b]]) for b in range(6)]
    assert texts2 == list(g["texts"])


@pytest.mark.parametrize("prec", [PREC_FP32, PREC_FP16])
@pytest.mark.parametrize("shape", [(1, 16), (3, 173), (5, 320), (2, 1081)])
def test_rec_vs_oracle_seeded(prec, shape):
    n, w = shape
    x = np.random.default_rng(w).standard_normal((n, 3, 48, w)).astype(np.float32)
    logits = nets.rec_logits(x)
    e = rec_engine(prec)
    out = e.infer_f32(x)
    T = logits.shape[1]
    assert e.tokens(w) == T and out["ids"].shape == (n, T)
    _check_rec(out, logits, prec, T)


def test_rec_softmax_compat_output(golden_dir):
    g = np.load(os.path.join(golden_dir, "rec_randn_3x48x173.npz"))
    out = rec_engine(PREC_FP32).infer_f32(g["x"], want_softmax=True)
    sm = out["softmax"]
    want = nets.rec_forward(g["x"])
    assert sm.shape == want.shape
    assert np.abs(sm - want).max() <= 1e-5
    assert np.array_equal(sm.argmax(2), g["ids"])
    res = P.ctc_decode(sm, nets.load_characters())
    assert [t for t, _ in res] == [t for t, _ in P.ctc_decode(want, nets.load_characters())]


def test_rec_chunking_is_invisible():
    x = np.random.default_rng(9).standard_normal((7, 3, 48, 96)).astype(np.float32)
    e = rec_engine(PREC_FP32)
    a = e.infer_f32(x)
    e.set_chunk_crops(3)
    b = e.infer_f32(x)
    e.set_chunk_crops(512)
    for k in ("ids", "probs", "text_ids", "text_len", "conf"):
        assert np.array_equal(a[k], b[k]), k


def test_device_pointer_path_matches_host_path():
    import torch
    rng = np.random.default_rng(1)
    pages = rng.integers(0, 256, (2, 96, 128, 3), dtype=np.uint8)
    e = det_engine(PREC_FP16)
    p_host, b_host = e.infer_u8(pages)
    d = torch.from_numpy(pages).cuda()
    p_dev, b_dev = e.infer_u8(d, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(p_dev.cpu().numpy(), p_host) and np.array_equal(b_dev.cpu().numpy(), b_host)
