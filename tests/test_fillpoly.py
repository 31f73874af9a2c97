"""oracle/fillpoly.py (restatement of cv2.fillPoly + DBPostProcess.box_score_fast, the oracle of the next-round GPU box scorer)
pinned against cv2 itself."""
import cv2
import numpy as np

from oracle import fillpoly


def _quads(seed, n, inside=True):
    rng = np.random.default_rng(seed)
    for t in range(n):
        h, w = int(rng.integers(5, 90)), int(rng.integers(5, 200))
        if t % 3 == 0:       # rotated rectangles, like DB boxes
            cx, cy = rng.uniform(0, w), rng.uniform(0, h)
            bw, bh, a = rng.uniform(2, w), rng.uniform(2, h / 1.5), rng.uniform(-0.6, 0.6)
            c, s = np.cos(a), np.sin(a)
            pts = np.array([[cx + x * c - y * s, cy + x * s + y * c] for x, y in ((-bw / 2, -bh / 2), (bw / 2, -bh / 2), (bw / 2, bh / 2), (-bw / 2, bh / 2))])
        else:                # arbitrary (also self-intersecting) quads
            pts = np.stack([rng.uniform(-5, w + 5, 4), rng.uniform(-5, h + 5, 4)], 1)
        if inside:
            pts = np.clip(pts, [0, 0], [w - 1, h - 1])
        yield h, w, pts.astype(np.int32)


def test_fill_poly_bit_identical_to_cv2_for_polygons_inside_the_mask():
    for h, w, p in _quads(0, 300):
        ref = np.zeros((h, w), np.uint8)
        cv2.fillPoly(ref, p.reshape(1, -1, 2), 1)
        assert np.array_equal(fillpoly.fill_poly(h, w, p), ref), (h, w, p.tolist())


def test_fill_poly_outside_vertices_mostly_identical():
    bad = 0
    for h, w, p in _quads(1, 150, inside=False):
        ref = np.zeros((h, w), np.uint8)
        cv2.fillPoly(ref, p.reshape(1, -1, 2), 1)
        got = fillpoly.fill_poly(h, w, p)
        bad += not np.array_equal(got, ref)
    assert bad <= 12        # the misses sit on clipped border columns / degenerate slivers (see the module docstring)


def test_box_score_fast_equals_cv2_mean_over_fillpoly_mask():
    rng = np.random.default_rng(3)
    prob = rng.random((160, 240)).astype(np.float32)
    for t in range(60):
        cx, cy = rng.uniform(10, 230), rng.uniform(10, 150)
        bw, bh, a = rng.uniform(6, 120), rng.uniform(4, 40), rng.uniform(-0.5, 0.5)
        c, s = np.cos(a), np.sin(a)
        box = np.float32([[cx + x * c - y * s, cy + x * s + y * c] for x, y in ((-bw / 2, -bh / 2), (bw / 2, -bh / 2), (bw / 2, bh / 2), (-bw / 2, bh / 2))])
        box = np.clip(box, [0, 0], [239.4, 159.4]).astype(np.float32)
        h, w = prob.shape
        b = box.copy()
        xmin = int(np.clip(np.floor(b[:, 0].min()), 0, w - 1)); xmax = int(np.clip(np.ceil(b[:, 0].max()), 0, w - 1))
        ymin = int(np.clip(np.floor(b[:, 1].min()), 0, h - 1)); ymax = int(np.clip(np.ceil(b[:, 1].max()), 0, h - 1))
        mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), np.uint8)
        b[:, 0] -= xmin; b[:, 1] -= ymin
        cv2.fillPoly(mask, b.reshape(1, -1, 2).astype(np.int32), 1)
        want = cv2.mean(prob[ymin:ymax + 1, xmin:xmax + 1], mask)[0]
        assert abs(fillpoly.box_score_fast(prob, box) - want) <= 1e-9
