"""GPU end-to-end test of the facade layer (rapiddoc_b200.ocr.B200OcrModel, the RapidOcrModel mirror) on a
real page: detection boxes vs the reference-net boxes, recognised text vs the reference flow's text."""
import os

import cv2
import numpy as np
import pytest

from rapiddoc_b200 import PREC_FP16, PREC_FP32
from rapiddoc_b200.ocr import B200OcrModel

pytestmark = pytest.mark.gpu


def _page(golden_dir):
    g = np.load(os.path.join(golden_dir, "page_img5_e2e.npz"))
    return cv2.imdecode(g["png"], cv2.IMREAD_COLOR), g


@pytest.mark.parametrize("prec", [PREC_FP32, PREC_FP16])
def test_det_then_rec_on_real_page(prec, golden_dir):
    img, g = _page(golden_dir)
    model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, enable_merge_det_boxes=False, precision=prec)
    # det only: [(boxes, elapse)] like RapidOcrModel.det_batch_predict
    (boxes, _), = model.det_batch_predict([img], max_batch_size=1)
    want = g["boxes"]
    assert len(boxes) == len(want)
    tol = 0 if prec == PREC_FP32 else 2          # fp16: box corners may move by a pixel or two
    assert np.abs(np.asarray(boxes, np.float32) - want).max() <= tol
    # det + rec
    res = model.ocr(img, det=True, rec=True)[0]
    texts = [r[1][0] for r in res]
    want_texts = list(g["texts"])
    if prec == PREC_FP32:
        assert texts == want_texts
        assert np.abs(np.array([r[1][1] for r in res]) - g["conf"]).max() <= 2e-4
    else:
        same = sum(a == b for a, b in zip(texts, want_texts))
        assert len(texts) == len(want_texts) and same >= len(want_texts) - 1, (texts, want_texts)
    # rec only on caller-supplied crops (the _run_ocr_rec_postprocess call shape)
    crops = [img[int(b[0][1]):int(b[2][1]), int(b[0][0]):int(b[2][0])] for b in want[:4]]
    out = model.ocr(crops, det=False, rec=True)[0]
    assert len(out) == 4 and all(isinstance(t, str) and 0.0 <= s <= 1.0 for t, s in out)


def test_custom_plugin_batch_predict(golden_dir):
    from rapiddoc_b200.plugin import B200OcrCustomModel
    img, g = _page(golden_dir)
    plug = B200OcrCustomModel(device=0, precision=PREC_FP16)
    out = plug.batch_predict([img, img[:200]])
    assert len(out) == 2 and "Chapter 3" in out[0] and isinstance(out[1], str)
