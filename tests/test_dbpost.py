"""Host half of the window-level DBPostProcess (rapiddoc_b200/dbpost.py) against the per-box oracle restatement
(oracle/ocr_post.py) and against OpenCV itself; the closed-form fillPoly raster of the GPU box scorer against cv2.fillPoly
(through the host build of the same function, rdb_debug_fill_quad).  No GPU needed."""
import os

import cv2
import numpy as np

from oracle import ocr_post as P
from rapiddoc_b200 import _lib, dbpost


def _quads(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        cx, cy = rng.uniform(20, 900), rng.uniform(20, 900)
        bw, bh, a = rng.uniform(3, 600), rng.uniform(3, 80), rng.uniform(-1.6, 1.6)
        c, s = np.cos(a), np.sin(a)
        out.append([[cx + x * c - y * s, cy + x * s + y * c] for x, y in ((-bw / 2, -bh / 2), (bw / 2, -bh / 2), (bw / 2, bh / 2), (-bw / 2, bh / 2))])
    return np.array(out, np.float32)


def test_area_and_perimeter_bit_identical_to_cv2():
    q = _quads(0, 4000)
    area = dbpost.contour_area_f32(q)
    length = dbpost.arc_length_f32(q)
    for i in range(len(q)):
        assert area[i] == cv2.contourArea(q[i])
        assert length[i] == cv2.arcLength(q[i], True)


def test_mini_boxes_ordering_equals_oracle():
    rng = np.random.default_rng(1)
    for _ in range(300):
        pts = (rng.uniform(0, 200, (int(rng.integers(3, 30)), 1, 2))).astype(np.int32)
        want, _ = P.get_mini_boxes(pts)
        got = dbpost.mini_boxes(cv2.boxPoints(cv2.minAreaRect(pts))[None])[0]
        assert np.array_equal(got, want)
    # axis-aligned rectangles: ties in x are resolved by the stable sort, as Python's sorted() does
    for _ in range(100):
        x0, y0 = rng.integers(0, 100, 2)
        w, h = rng.integers(3, 80, 2)
        pts = np.array([[[x0, y0]], [[x0 + w, y0]], [[x0 + w, y0 + h]], [[x0, y0 + h]]], np.int32)
        want, _ = P.get_mini_boxes(pts)
        got = dbpost.mini_boxes(cv2.boxPoints(cv2.minAreaRect(pts))[None])[0]
        assert np.array_equal(got, want)


def _synthetic_prob(seed, h=512, w=640, lines=18):
    """A DB-like probability map: blurred rotated rectangles of text-line shape, some touching the page border."""
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), np.float32)
    for k in range(lines):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        bw, bh, a = rng.uniform(20, 300), rng.uniform(6, 30), rng.uniform(-0.3, 0.3) if k % 4 else rng.uniform(-1.5, 1.5)
        box = cv2.boxPoints(((cx, cy), (bw, bh), np.degrees(a)))
        cv2.fillPoly(m, [box.astype(np.int32)], float(rng.uniform(0.5, 1.0)))
    m = cv2.GaussianBlur(m, (0, 0), 1.5)
    m += rng.uniform(0, 0.05, m.shape).astype(np.float32)
    return np.clip(m, 0, 1).astype(np.float32)


def test_window_boxes_identical_to_per_box_oracle():
    probs = np.stack([_synthetic_prob(s) for s in range(6)])
    bitmaps = np.stack([P.db_bitmap(p, 0.3, True) for p in probs])
    for box_thresh, ratio, dest in ((0.3, 1.8, (512, 640)), (0.5, 1.6, (700, 900))):
        got = dbpost.window_boxes(bitmaps, [dest] * len(probs), dbpost.cv2_score_fn(probs), box_thresh, ratio)
        total = 0
        for i in range(len(probs)):
            wb, ws = P.db_postprocess(probs[i][None, None], dest, 0.3, box_thresh, ratio, True)
            gb, gs = got[i]
            assert len(gb) == len(wb)
            total += len(wb)
            if len(wb):
                assert np.array_equal(np.asarray(gb), np.asarray(wb))
                assert np.array_equal(np.asarray(gs), np.asarray(ws))
        assert total > 30


def test_window_boxes_on_the_reference_page(golden_dir):
    g = np.load(os.path.join(golden_dir, "det_page_img5.npz"))
    prob = g["prob"][0]
    shape = tuple(int(v) for v in g["shape"])
    bm = P.db_bitmap(prob[0], 0.3, True)[None]
    (boxes, scores), = dbpost.window_boxes(bm, [shape], dbpost.cv2_score_fn(prob), 0.3, 1.8)
    boxes = np.array(P.sorted_boxes(boxes))
    assert np.array_equal(boxes, g["boxes"])


def test_closed_form_fill_quad_equals_cv2_fillpoly():
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for t in range(3000):
        h, w = int(rng.integers(3, 90)), int(rng.integers(3, 300))
        if t % 3 == 0:
            cx, cy = rng.uniform(0, w), rng.uniform(0, h)
            bw, bh, a = rng.uniform(2, w), rng.uniform(2, h / 1.5), rng.uniform(-0.6, 0.6)
            c, s = np.cos(a), np.sin(a)
            pts = np.array([[cx + x * c - y * s, cy + x * s + y * c] for x, y in ((-bw / 2, -bh / 2), (bw / 2, -bh / 2), (bw / 2, bh / 2), (-bw / 2, bh / 2))])
        else:
            pts = np.stack([rng.uniform(-5, w + 5, 4), rng.uniform(-5, h + 5, 4)], 1)
        pts = np.ascontiguousarray(np.clip(pts, [0, 0], [w - 1, h - 1]).astype(np.int32))
        ref = np.zeros((h, w), np.uint8)
        cv2.fillPoly(ref, pts.reshape(1, -1, 2), 1)
        got = np.zeros((h, w), np.uint8)
        assert lib.rdb_debug_fill_quad(pts.ctypes.data, w, h, got.ctypes.data) == 0
        assert np.array_equal(ref, got), (h, w, pts.tolist())


def test_clipper_small_delta_and_near_collinear():
    """OffsetPoint's early return must not advance k — product (C++) and oracle agree on thin / tiny-delta quads."""
    rng = np.random.default_rng(9)
    for _ in range(400):
        x0, y0 = rng.integers(0, 50, 2)
        w, h = int(rng.integers(3, 200)), int(rng.integers(1, 4))
        sk = int(rng.integers(0, 2))
        box = np.array([[x0, y0], [x0 + w, y0 + sk], [x0 + w, y0 + sk + h], [x0, y0 + h]], np.float32)
        d = float(rng.uniform(0.2, 3.0))
        want = P.clipper_offset_round(box, d)
        got = dbpost.clipper_offset(box, d).reshape(-1, 2)
        assert np.array_equal(got, want)
