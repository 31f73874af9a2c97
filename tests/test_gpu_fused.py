"""The fused tcgen05 kernels (stem_planar, head_planar, mlp_tc, stem1_tc, packed-half dw7x7, fast GELU) against their
unfused / SIMT predecessors, which stay selectable through environment switches read per call, and against the CPU
oracle — on shapes whose tiles are partial in both directions.  fp16 mode only (the fp32 mode never takes these paths)."""
import os

import numpy as np
import pytest

from rapiddoc_b200 import _lib

from oracle import nets, ocr_post as P
from rapiddoc_b200 import PREC_FP16
from rapiddoc_b200.engine import DetEngine, RecEngine

pytestmark = pytest.mark.gpu

SWITCHES = [("RDB_STEM", "unfused"), ("RDB_STEM", "copy"), ("RDB_HEAD", "unfused"), ("RDB_MLP", "unfused"), ("RDB_STEM1", "simt"),
            ("RDB_DW7", "f32"), ("RDB_GELU", "exact")]


def _pages(n, h, w, seed):
    rng = np.random.default_rng(seed)
    pages = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    pages[:, h // 5: h // 2, w // 7: w // 2] = 250
    pages[:, h // 2:, :] = (pages[:, h // 2:, :] // 32) * 32
    return pages


@pytest.fixture(scope="module")
def det():
    return DetEngine(device=0, precision=PREC_FP16)


@pytest.mark.parametrize("shape", [(2, 288, 352), (1, 416, 608), (3, 64, 32)])
def test_det_fused_paths_match_unfused_and_oracle(det, shape):
    n, h, w = shape
    pages = _pages(n, h, w, seed=h + w)
    want = np.stack([nets.det_forward(P.det_preprocess(p, limit_side_len=4096))[0, 0] for p in pages])
    prob, bm = det.infer_u8(pages, thresh=0.3, use_dilation=True)
    assert np.abs(prob - want).max() <= 3e-2
    assert np.array_equal(bm, np.stack([P.db_bitmap(p, 0.3, True) for p in prob]))
    for name, val in SWITCHES:
        os.environ[name] = val
        _lib.load().rdb_switches_reload()        # switches are cached per process
        try:
            alt, alt_bm = det.infer_u8(pages, thresh=0.3, use_dilation=True)
        finally:
            del os.environ[name]
            _lib.load().rdb_switches_reload()
        assert np.abs(alt - want).max() <= 3e-2, (name, val)
        assert np.abs(alt - prob).max() <= 2e-2, (name, val)          # same math, different rounding points
        flips = (alt > 0.3) != (prob > 0.3)
        assert not (flips & (np.abs(want - 0.3) > 2e-2)).any(), (name, val)


def test_det_fused_deterministic_and_batch_invariant(det):
    pages = _pages(5, 160, 224, seed=7)
    a, _ = det.infer_u8(pages)
    b, _ = det.infer_u8(pages)
    assert np.array_equal(a, b)
    c, _ = det.infer_u8(pages[2:3])
    assert np.array_equal(a[2:3], c)        # a page's result does not depend on its neighbours in the batch


def test_rec_fused_paths_match_unfused():
    rec = RecEngine(device=0, precision=PREC_FP16)
    rng = np.random.default_rng(11)
    x = rng.standard_normal((5, 3, 48, 200)).astype(np.float32)
    base = rec.infer_f32(x)
    logits = nets.rec_logits(x)
    srt = np.sort(logits, axis=2)
    margin = srt[:, :, -1] - srt[:, :, -2]
    assert not ((base["ids"] != logits.argmax(2)) & (margin > 0.25)).any()
    for name, val in [("RDB_MLP", "unfused"), ("RDB_STEM1", "simt"), ("RDB_GELU", "exact")]:
        os.environ[name] = val
        _lib.load().rdb_switches_reload()
        try:
            alt = rec.infer_f32(x)
        finally:
            del os.environ[name]
            _lib.load().rdb_switches_reload()
        assert not ((alt["ids"] != base["ids"]) & (margin > 0.25)).any(), (name, val)
