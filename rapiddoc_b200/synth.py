"""Deterministic synthetic inputs for BASELINE.json's configs:
text rendered with cv2.putText so the DB post-process / CTC decode see realistic blobs.
CPU only, numpy.random.default_rng(seed)."""
import string

import cv2
import numpy as np

_ALPHA = string.ascii_letters + string.digits + "  .,;:-()%/"
_FONTS = [cv2.FONT_HERSHEY_SIMPLEX, cv2.FONT_HERSHEY_DUPLEX, cv2.FONT_HERSHEY_COMPLEX, cv2.FONT_HERSHEY_TRIPLEX]


def _rand_text(rng, n):
    return "".join(_ALPHA[i] for i in rng.integers(0, len(_ALPHA), n))


def det_page(rng, h=1024, w=1024, lines=60):
    """One BGR uint8 page: near-white background, ~`lines` text lines, glyph height 16-40 px."""
    img = np.full((h, w, 3), 255, np.uint8)
    img[:] = rng.integers(235, 256, 3, dtype=np.uint8)
    y = int(rng.integers(20, 50))
    count = 0
    while y < h - 20 and count < lines:
        gh = int(rng.integers(16, 41))
        font = _FONTS[int(rng.integers(0, len(_FONTS)))]
        scale = gh / 22.0
        thick = 1 if gh < 26 else 2
        x = int(rng.integers(10, max(11, w // 6)))
        nchar = int(rng.integers(8, max(9, int((w - x) / (gh * 0.62)))))
        txt = _rand_text(rng, nchar)
        color = tuple(int(c) for c in rng.integers(0, 90, 3))
        cv2.putText(img, txt, (x, y + gh), font, scale, color, thick, cv2.LINE_AA)
        y += gh + int(rng.integers(4, max(5, gh // 2 + 5)))
        count += 1
    return img


def det_pages(n, h=1024, w=1024, seed=1, lines=60):
    rng = np.random.default_rng(seed)
    return np.stack([det_page(rng, h, w, lines) for _ in range(n)])


def rec_crop(rng, h=48, w=320):
    img = np.full((h, w, 3), 255, np.uint8)
    img[:] = rng.integers(225, 256, 3, dtype=np.uint8)
    n = int(rng.integers(8, 29))
    txt = _rand_text(rng, n)
    font = _FONTS[int(rng.integers(0, len(_FONTS)))]
    scale = min(1.2, (w - 12) / (n * 19.0))
    color = tuple(int(c) for c in rng.integers(0, 90, 3))
    cv2.putText(img, txt, (6, int(h * 0.72)), font, scale, color, 2 if scale > 0.9 else 1, cv2.LINE_AA)
    return img


def rec_crops(n, h=48, w=320, seed=2):
    rng = np.random.default_rng(seed)
    return np.stack([rec_crop(rng, h, w) for _ in range(n)])
