"""GPU tests at BASELINE.json's FULL sizes (det: 64 pages of 1024x1024, rec: 512 crops of 48x320).  The CPU oracle
is too slow for the whole batch, so the batch is checked through size-independent properties plus an oracle spot
check on a few sampled units:
  * run-to-run determinism (bit-identical outputs),
  * batch-composition invariance (a unit's result does not depend on its neighbours / on internal chunking / lanes),
  * oracle parity on sampled units.
"""
import numpy as np
import pytest

from oracle import nets, ocr_post as P
from rapiddoc_b200 import PREC_FP16, synth
from rapiddoc_b200.engine import DetEngine, RecEngine

pytestmark = pytest.mark.gpu


def test_det_full_batch_properties():
    base = synth.det_pages(8, 1024, 1024, seed=1)
    pages = np.stack([np.roll(base[i % 8], shift=(7 * (i // 8), 13 * (i // 8)), axis=(0, 1)) for i in range(64)])
    e = DetEngine(0, PREC_FP16)
    prob, bm = e.infer_u8(pages)
    prob2, bm2 = e.infer_u8(pages)
    assert np.array_equal(prob, prob2) and np.array_equal(bm, bm2), "not deterministic"
    # batch-composition invariance: pages 5, 37, 63 alone == inside the batch (different chunk / lane / position)
    for i in (5, 37, 63):
        p1, b1 = e.infer_u8(pages[i:i + 1])
        assert np.array_equal(p1[0], prob[i]) and np.array_equal(b1[0], bm[i]), f"page {i} depends on its batch"
    # bitmap == cv2 binarise+dilate of our own prob map, for every page
    for i in range(0, 64, 9):
        assert np.array_equal(bm[i], P.db_bitmap(prob[i], 0.3, True))
    # oracle spot check (2 pages)
    for i in (0, 41):
        want = nets.det_forward(P.det_preprocess(pages[i], limit_side_len=1024))[0, 0]
        d = np.abs(prob[i] - want)
        assert d.max() <= 3e-2
        flips = (prob[i] > 0.3) != (want > 0.3)
        assert not (flips & (np.abs(want - 0.3) > 2e-2)).any()
        # the text blobs the detector exists for: DB boxes from our maps == boxes from the oracle maps (+-2 px)
        ours = P.boxes_from_bitmap(prob[i], bm[i], 1024, 1024, box_thresh=0.5, unclip_ratio=1.6)[0]
        ref = P.boxes_from_bitmap(want, P.db_bitmap(want, 0.3, True), 1024, 1024, box_thresh=0.5, unclip_ratio=1.6)[0]
        assert len(ours) == len(ref) and len(ref) > 20
        assert np.abs(ours.astype(np.int64) - ref.astype(np.int64)).max() <= 2


def test_rec_full_batch_properties():
    crops = synth.rec_crops(512, 48, 320, seed=2)
    vw = np.full(512, 320, np.int32)
    e = RecEngine(0, PREC_FP16)
    out = e.infer_u8(crops, vw)
    out2 = e.infer_u8(crops, vw)
    for k in ("ids", "probs", "text_ids", "text_len", "conf"):
        assert np.array_equal(out[k], out2[k]), f"{k} not deterministic"
    sub = e.infer_u8(crops[100:116], vw[100:116])                 # same crops in a different batch
    assert np.array_equal(sub["ids"], out["ids"][100:116]) and np.array_equal(sub["probs"], out["probs"][100:116])
    # oracle spot check on 16 crops: decoded text must be identical, confidence close
    idx = list(range(0, 512, 32))
    x = np.stack([P.resize_norm_img(crops[i], 320 / 48) for i in idx])
    want = P.ctc_decode(nets.rec_forward(x), nets.load_characters())
    chars = nets.load_characters()
    for j, i in enumerate(idx):
        text = "".join(chars[t] for t in out["text_ids"][i][: out["text_len"][i]])
        assert text == want[j][0], (i, text, want[j][0])
        assert abs(float(out["conf"][i]) - want[j][1]) <= 2e-2
    assert sum(len(w[0]) > 0 for w in want) >= 12            # the synthetic lines are readable text
