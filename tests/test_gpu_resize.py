"""GPU: DetPreProcess' cv2.resize(INTER_LINEAR, uint8) re-implemented as a kernel must be BIT-EXACT with OpenCV,
standalone and fused in front of the detector."""
import numpy as np
import cv2
import pytest

from oracle import nets, ocr_post as P
from rapiddoc_b200 import PREC_FP32, _lib
from rapiddoc_b200.engine import DetEngine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(720, 960, 704, 960), (2339, 1654, 960, 672), (100, 300, 64, 192), (1500, 1000, 960, 640),
                                   (333, 517, 320, 512), (64, 64, 64, 64), (50, 31, 96, 64), (30, 40, 33, 47), (37, 211, 48, 274), (3500, 2000, 960, 544)])
def test_resize_bit_exact_vs_cv2(shape):
    h, w, dh, dw = shape
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    img[1, : h // 2] = (img[1, : h // 2] // 128) * 255        # hard edges
    out = np.empty((2, dh, dw, 3), np.uint8)
    _lib.check(_lib.load().rdb_resize_linear_u8(0, img.ctypes.data, 2, h, w, out.ctypes.data, dh, dw, None))
    for i in range(2):
        assert np.array_equal(out[i], cv2.resize(img[i], (dw, dh)))


def test_det_with_gpu_resize_matches_host_resize_path():
    rng = np.random.default_rng(3)
    page = rng.integers(0, 256, (1, 1100, 1500, 3), dtype=np.uint8)
    page[:, 300:500, 200:1200] = 255
    x = P.det_preprocess(page[0], limit_side_len=960)                 # oracle: cv2.resize + normalise on the CPU
    rh, rw = x.shape[2:]
    e = DetEngine(0, PREC_FP32)
    prob_gpu_resize, bm1 = e.infer_u8(page, resize_to=(rh, rw))
    prob_host_resize, bm2 = e.infer_u8(cv2.resize(page[0], (rw, rh))[None])
    assert np.array_equal(prob_gpu_resize, prob_host_resize) and np.array_equal(bm1, bm2)
    want = nets.det_forward(x)[0, 0]
    assert np.abs(prob_gpu_resize[0] - want).max() <= 2e-5
