"""rdb_contours_* (native Suzuki-Abe border following on host threads) pinned against cv2.findContours itself:
RETR_LIST + CHAIN_APPROX_SIMPLE — same number of contours, same order, same points."""
import cv2
import numpy as np

from rapiddoc_b200 import dbpost, synth


def _check(bitmaps):
    per_page, sizes, pts = dbpost.find_contours_window(bitmaps)
    so = po = 0
    total = 0
    for i, bm in enumerate(bitmaps):
        res = cv2.findContours(bm, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        want = res[0] if len(res) == 2 else res[1]
        assert per_page[i] == len(want), (i, per_page[i], len(want))
        for c in want:
            n = int(sizes[so])
            got = pts[po: po + n]
            assert n == len(c) and np.array_equal(got, c), (i, so, got.reshape(-1, 2)[:6].tolist(), c.reshape(-1, 2)[:6].tolist())
            so += 1
            po += n
        total += len(want)
    return total


def test_random_noise_and_blobs():
    rng = np.random.default_rng(0)
    maps = []
    for k in range(40):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 90))
        maps.append((rng.random((h, w)) < rng.uniform(0.05, 0.95)).astype(np.uint8))
    for bm in maps:                      # different sizes: one window each
        _check(bm[None])
    big = np.stack([(cv2.GaussianBlur(rng.random((200, 300)).astype(np.float32), (0, 0), s) > t).astype(np.uint8)
                    for s, t in ((1.0, 0.5), (2.0, 0.5), (3.0, 0.49), (5.0, 0.5), (1.5, 0.45), (0.7, 0.55))])
    assert _check(big) > 200


def test_edge_cases():
    z = np.zeros((5, 7), np.uint8)
    assert _check(z[None]) == 0
    o = np.ones((5, 7), np.uint8)
    assert _check(o[None]) == 1
    single = z.copy(); single[2, 3] = 255
    ring = np.ones((9, 9), np.uint8); ring[3:6, 3:6] = 0
    diag = np.eye(9, dtype=np.uint8)
    border = z.copy(); border[0, :] = 1; border[:, 0] = 1; border[-1, -1] = 1
    checker = (np.indices((9, 9)).sum(0) % 2).astype(np.uint8)
    for bm in (single, border):
        _check(bm[None])
    for bm in (ring, diag, checker):
        _check(bm[None])


def test_db_like_text_bitmaps():
    """Dilated text-line masks of the synthetic pages (what boxes_from_bitmap sees), 1024x1024."""
    pages = synth.det_pages(3, 1024, 1024, seed=2)
    bms = []
    for p in pages:
        g = cv2.cvtColor(p, cv2.COLOR_BGR2GRAY)
        m = (cv2.blur(255 - g, (15, 7)) > 25).astype(np.uint8)
        bms.append(cv2.dilate(m, np.array([[1, 1], [1, 1]], np.uint8)))
    assert _check(np.stack(bms)) > 40
