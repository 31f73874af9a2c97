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


def seal_image(seed=0, size=320, text="RAPIDDOCSEALTEXT2026"):
    """A synthetic round seal (BGR uint8): red ring, glyphs set along an arc, one straight line in the middle — the curved-text
    input of the seal detector (pp-ocrv4_mobile_seal_det.onnx)."""
    import cv2
    rng = np.random.RandomState(seed)
    img = np.full((size, size, 3), 255, np.uint8)
    c = size // 2
    cv2.circle(img, (c, c), int(size * 0.45), (0, 0, 255), 3)
    start = 1.15 + 0.1 * rng.rand()
    for k, ch in enumerate(text):
        ang = np.pi * (start + 1.4 * k / len(text))
        r = size * 0.36
        x, y = int(c + r * np.cos(ang)), int(c + r * np.sin(ang))
        glyph = np.full((40, 40, 3), 255, np.uint8)
        cv2.putText(glyph, ch, (8, 30), cv2.FONT_HERSHEY_SIMPLEX, 1.0, (0, 0, 255), 2, cv2.LINE_AA)
        M = cv2.getRotationMatrix2D((20, 20), -np.degrees(ang) - 90, 1.0)
        glyph = cv2.warpAffine(glyph, M, (40, 40), borderValue=(255, 255, 255))
        y0, x0 = y - 20, x - 20
        if 0 <= y0 and y0 + 40 <= size and 0 <= x0 and x0 + 40 <= size:
            img[y0:y0 + 40, x0:x0 + 40] = np.minimum(img[y0:y0 + 40, x0:x0 + 40], glyph)
    cv2.putText(img, "CONTRACT SEAL", (c - 90, c + 10), cv2.FONT_HERSHEY_SIMPLEX, 0.8, (0, 0, 255), 2, cv2.LINE_AA)
    return img


def table_image(seed=0, rows=4, cols=3, h=300, w=400, lines=True):
    """A synthetic table crop (BGR uint8): a ruled (or borderless) grid with short cell texts — the input of the table
    structure model (SLANet)."""
    import cv2
    rng = np.random.RandomState(seed)
    img = np.full((h, w, 3), 255, np.uint8)
    x0, y0, x1, y1 = 10, 20, w - 10, h - 40
    ch, cw = (y1 - y0) // rows, (x1 - x0) // cols
    if lines:
        for r in range(rows + 1):
            cv2.line(img, (x0, y0 + r * ch), (x0 + cols * cw, y0 + r * ch), (0, 0, 0), 1)
        for c in range(cols + 1):
            cv2.line(img, (x0 + c * cw, y0), (x0 + c * cw, y0 + rows * ch), (0, 0, 0), 1)
    for r in range(rows):
        for c in range(cols):
            txt = "".join(chr(ord("a") + int(v)) for v in rng.randint(0, 26, 3)) + str(int(rng.randint(0, 100)))
            cv2.putText(img, txt, (x0 + c * cw + 8, y0 + r * ch + int(ch * 0.65)), cv2.FONT_HERSHEY_SIMPLEX, 0.6, (0, 0, 0), 2)
    return img
