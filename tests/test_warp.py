"""get_rotate_crop_image / cv2.warpPerspective(INTER_CUBIC, BORDER_REPLICATE) (rapid_doc/utils/ocr_utils.py:494-537).
CPU: the numpy restatement (oracle/warp.py) is pinned bit-for-bit against cv2 itself, and the C++ weight table the CUDA kernel
uses equals the restated one.  GPU: rdb_warp_crops against cv2 on the same quads (bit-exact, including rot90 and quads that
leave the page)."""
import ctypes as C

import cv2
import numpy as np
import pytest

from oracle import warp
from rapiddoc_b200 import _lib


def _page(seed=0, h=300, w=400):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[50:120, 60:300] = cv2.GaussianBlur(img[50:120, 60:300], (5, 5), 0)
    cv2.putText(img, "B200 warp", (30, 200), cv2.FONT_HERSHEY_SIMPLEX, 1.4, (10, 10, 10), 3)
    return img


def _quads(seed, n, h=300, w=400):
    rng = np.random.default_rng(seed)
    out = []
    for t in range(n):
        cx, cy = rng.uniform(40, w - 40), rng.uniform(30, h - 30)
        bw, bh = rng.uniform(20, 260), rng.uniform(8, 60)
        if t % 7 == 0:
            bw, bh = bh, bw * 0.6 + 10                      # tall box -> rot90 branch
        a = rng.uniform(-0.4, 0.4)
        c, s = np.cos(a), np.sin(a)
        p = np.float32([[cx + x * c - y * s, cy + x * s + y * c] for x, y in ((-bw / 2, -bh / 2), (bw / 2, -bh / 2), (bw / 2, bh / 2), (-bw / 2, bh / 2))])
        p += rng.uniform(-3, 3, p.shape).astype(np.float32)
        if t % 5 == 0:
            p = np.round(p)                                 # integer corners, as DBPostProcess produces
        if t % 11 == 0:
            p[:, 0] -= 60                                   # partly outside the page: replicate border
        out.append(p)
    return out


def _cv2_crop(img, p):
    cw = int(max(np.linalg.norm(p[0] - p[1]), np.linalg.norm(p[2] - p[3])))
    ch = int(max(np.linalg.norm(p[0] - p[3]), np.linalg.norm(p[1] - p[2])))
    std = np.float32([[0, 0], [cw, 0], [cw, ch], [0, ch]])
    dst = cv2.warpPerspective(img, cv2.getPerspectiveTransform(p, std), (cw, ch), borderMode=cv2.BORDER_REPLICATE, flags=cv2.INTER_CUBIC)
    return np.rot90(dst) if dst.shape[0] * 1.0 / dst.shape[1] >= 2 else dst


def test_oracle_warp_is_bit_identical_to_cv2():
    img = _page()
    for p in _quads(1, 24):
        assert np.array_equal(warp.get_rotate_crop_image(img, p), _cv2_crop(img, p))


def test_library_cubic_table_equals_restatement():
    lib = _lib.load()
    tab = np.zeros(1024 * 16, np.int16)
    _lib.check(lib.rdb_debug_cubic_tab(tab.ctypes.data_as(C.c_void_p)))
    want = warp.cubic_tab_2d().reshape(-1)
    assert np.array_equal(tab, want)
    assert (want.reshape(1024, 16).astype(np.int64).sum(1) == 32768).all()


@pytest.mark.gpu
def test_gpu_warp_crops_bit_exact_vs_cv2():
    from rapiddoc_b200.ocr import get_rotate_crop_images_gpu
    img = _page(3, 640, 900)
    quads = _quads(4, 60, 640, 900)
    quads.append(np.float32([[5, 5], [5, 5], [5, 5], [5, 5]]))      # degenerate: None
    got = get_rotate_crop_images_gpu(img, quads)
    assert got[-1] is None
    for g, p in zip(got[:-1], quads[:-1]):
        want = _cv2_crop(img, p)
        assert g.shape == want.shape
        assert np.array_equal(g, want)


@pytest.mark.gpu
def test_gpu_warp_device_page_and_large_crop():
    import torch
    from rapiddoc_b200.ocr import get_rotate_crop_images_gpu
    img = _page(5, 1200, 1600)
    quads = [np.float32([[100, 100], [1500, 130], [1495, 330], [95, 300]]), np.float32([[200, 400], [260, 402], [262, 1100], [198, 1098]])]
    got = get_rotate_crop_images_gpu(torch.from_numpy(img).cuda(), quads)
    for g, p in zip(got, quads):
        assert np.array_equal(g, _cv2_crop(img, p))


@pytest.mark.gpu
def test_gpu_resize_pack_bit_exact_vs_cv2():
    """resize_norm_img geometry: ragged crops -> [n,48,iw,3], each cv2.resize'd to its own width, zero to the right."""
    import math
    rng = np.random.default_rng(9)
    shapes = [(31, 200), (96, 412), (48, 320), (24, 77), (60, 15), (144, 300), (17, 640), (5, 9)]   # incl. exact 2x / 3x / identity height
    crops = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    ih = 48
    mx = max([320 / 48] + [w / float(h) for h, w in shapes])
    iw = int(ih * mx)
    vw = np.array([iw if math.ceil(ih * (w / float(h))) > iw else int(math.ceil(ih * (w / float(h)))) for h, w in shapes], np.int32)
    src = np.concatenate([c.reshape(-1) for c in crops])
    nb = np.array([c.size for c in crops], np.int64)
    offs = np.concatenate([[0], np.cumsum(nb)[:-1]]).astype(np.int64)
    sizes = np.array([[w, h] for h, w in shapes], np.int32)
    out = np.full((len(crops), ih, iw, 3), 7, np.uint8)
    _lib.check(_lib.load().rdb_resize_pack_u8(0, src.ctypes.data, int(src.size), len(crops), offs.ctypes.data, sizes.ctypes.data, vw.ctypes.data,
                                              out.ctypes.data, ih, iw, None))
    for i, c in enumerate(crops):
        assert np.array_equal(out[i, :, :vw[i]], cv2.resize(c, (int(vw[i]), ih))), shapes[i]
        assert not out[i, :, vw[i]:].any()


@pytest.mark.gpu
def test_facade_device_crops_equal_host_crops(golden_dir):
    """B200OcrModel with the crops kept on the GPU (warp + resize + pack on the device) returns exactly the boxes, texts and
    scores of the host cv2 path: both stages are bit-exact, so the recogniser sees identical bytes."""
    import os
    from rapiddoc_b200 import PREC_FP16
    from rapiddoc_b200.ocr import B200OcrModel
    g = np.load(os.path.join(golden_dir, "page_img5_e2e.npz"))
    img = cv2.imdecode(g["png"], cv2.IMREAD_COLOR)
    model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, enable_merge_det_boxes=False, precision=PREC_FP16)
    model.gpu_crop = True
    a = model.ocr(img, det=True, rec=True)[0]
    model.gpu_crop = False
    b = model.ocr(img, det=True, rec=True)[0]
    assert len(a) == len(b) and len(a) > 5
    for x, y in zip(a, b):
        assert x[0] == y[0] and x[1][0] == y[1][0] and x[1][1] == y[1][1]
