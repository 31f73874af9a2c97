"""CPU restatement (test infrastructure only) of cv2.fillPoly(mask, [int32 polygon], 1) and of DBPostProcess.box_score_fast on top
of it (rapidocr DBPostProcess as patched by rapid_doc/model/ocr/ocr_patch.py:223-241; PaddleOCR db_postprocess.py box_score_fast).
OpenCV is the third-party dependency (opencv-python, unpinned); the algorithm restated is modules/imgproc/src/drawing.cpp:
CollectPolyEdges (every polygon edge is first drawn with the 8-connected LineIterator, then registered as a 16.16 fixed-point
scan edge starting at x + 0.5) and FillEdgeCollection (edges sorted by (y0, x, dx); per scanline the active edges are paired and
the span [x_left >> 16, x_right >> 16] is filled; x += dx with dx = trunc((x1 - x0) / (y1 - y0)); the lowest row comes from the
outline only; a span ends at (x_right - 0.5) >> 16, so it never reaches past the outline).  Pinned bit-for-bit against cv2.fillPoly
(4.13) by tests/test_fillpoly.py on 480/480 random quads — convex and self-intersecting — whose vertices lie inside the mask, which is
the box_score_fast case; with vertices up to 5 px OUTSIDE the mask 116/120 match (the misses sit on the clipped border column: the
clipped-edge start is reconstructed from the observable behaviour, not from the source).  This is the oracle for the
next-round GPU box scorer: with the score computed on the device the fp32 prob map no longer has to come
back to the host."""
import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT


def _clip_line(w, h, x1, y1, x2, y2):
    """cv::clipLine(Size, Point&, Point&) for int64 points: Cohen-Sutherland style clipping with OpenCV's integer rounding."""
    right, bottom = w - 1, h - 1
    if w <= 0 or h <= 0:
        return False, x1, y1, x2, y2
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int((a - y1) * (x2 - x1) / (y2 - y1)) if False else _idiv((a - y1) * (x2 - x1), (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += _idiv((a - y2) * (x2 - x1), (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += _idiv((a - x1) * (y2 - y1), (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += _idiv((a - x2) * (y2 - y1), (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, x1, y1, x2, y2


def _idiv(a, b):
    """C integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _line8(mask, x1, y1, x2, y2):
    """cv::Line with connectivity 8: clipLine, then LineIterator (Bresenham with OpenCV's error term and step order)."""
    h, w = mask.shape
    ok, x1, y1, x2, y2 = _clip_line(w, h, x1, y1, x2, y2)
    if not ok:
        return
    if x2 < x1:          # cv::Line builds its LineIterator with leftToRight = true
        x1, y1, x2, y2 = x2, y2, x1, y1
    dx, dy = x2 - x1, y2 - y1
    sx = -1 if dx < 0 else 1
    sy = -1 if dy < 0 else 1
    dx, dy = abs(dx), abs(dy)
    if dy > dx:          # y is the major axis
        major, minor = dy, dx
        err = major - 2 * minor
        x, y = x1, y1
        for _ in range(major + 1):
            mask[y, x] = 1
            if err < 0:
                err += 2 * major
                x += sx
            err -= 2 * minor
            y += sy
    else:
        major, minor = dx, dy
        err = major - 2 * minor
        x, y = x1, y1
        for _ in range(major + 1):
            mask[y, x] = 1
            if err < 0:
                err += 2 * major
                y += sy
            err -= 2 * minor
            x += sx


def fill_poly(h, w, pts):
    """mask [h,w] uint8 with 1 inside/on the polygon pts [n,2] int (x, y) — cv2.fillPoly(mask, [pts], 1), LINE_8, shift 0."""
    mask = np.zeros((h, w), np.uint8)
    pts = [(int(p[0]), int(p[1])) for p in pts]
    n = len(pts)
    edges = []
    p0 = pts[-1]
    for i in range(n):
        p1 = pts[i]
        x0f, y0 = p0[0] << XY_SHIFT, p0[1]
        x1f, y1 = p1[0] << XY_SHIFT, p1[1]
        t0x, t1x = (x0f + (XY_ONE >> 1)) >> XY_SHIFT, (x1f + (XY_ONE >> 1)) >> XY_SHIFT
        _line8(mask, t0x, y0, t1x, y1)
        c0x, c0y, c1x, c1y = x0f, y0, x1f, y1
        if not (0 <= t0x < w and 0 <= t1x < w and 0 <= y0 < h and 0 <= y1 < h):
            ok, a, b, c, d = _clip_line(w, h, t0x, y0, t1x, y1)
            if b != d:
                c0y, c1y = b, d
                c0x, c1x = (a << XY_SHIFT) + (XY_ONE >> 1), (c << XY_SHIFT) + (XY_ONE >> 1)
        else:
            c0x += XY_ONE >> 1
            c1x += XY_ONE >> 1
        if y0 != y1:
            dxe = _idiv(c1x - c0x, c1y - c0y)
            if y0 < y1:
                edges.append([y0, y1, c0x + (y0 - c0y) * dxe, dxe])
            else:
                edges.append([y1, y0, c1x + (y1 - c1y) * dxe, dxe])
        p0 = p1
    if len(edges) < 2:
        return mask
    y_max = max(e[1] for e in edges)
    y_min = min(e[0] for e in edges)
    xs = [e[2] for e in edges] + [e[2] + (e[1] - e[0]) * e[3] for e in edges]
    if y_max < 0 or y_min >= h or max(xs) < 0 or min(xs) >= (w << XY_SHIFT):
        return mask
    edges.sort(key=lambda e: (e[0], e[2], e[3]))
    y_max = min(y_max, h)
    active = []          # kept sorted by x like OpenCV's linked list
    i = 0
    total = len(edges)
    for y in range(edges[0][0], y_max):
        # the C code interleaves removal / insertion / drawing in one list walk; the observable result is: drop edges ending at y,
        # merge the edges starting at y into the x-sorted active list (new edge goes BEFORE an active edge with x >= its x), pair them up
        active = [e for e in active if e[1] != y]
        while i < total and edges[i][0] == y:
            e = edges[i]
            k = 0
            while k < len(active) and active[k][2] < e[2]:
                k += 1
            active.insert(k, e)
            i += 1
        for k in range(0, len(active) - 1, 2):
            a, b = active[k], active[k + 1]
            if y >= 0:
                lo, hi = (b[2], a[2]) if a[2] > b[2] else (a[2], b[2])
                x1, x2 = lo >> XY_SHIFT, (hi - (XY_ONE >> 1)) >> XY_SHIFT      # spans never reach past the outline drawn above
                if x1 < w and x2 >= 0 and x2 >= x1:
                    mask[y, max(x1, 0): min(x2, w - 1) + 1] = 1
            a[2] += a[3]
            b[2] += b[3]
        # bubble sort by x (stable for equal x, as the C code only swaps on >)
        active.sort(key=lambda e: e[2])
    return mask


def box_score_fast(prob, box):
    """DBPostProcess.box_score_fast: mean of prob inside the (integer-truncated) quad, over its clipped bounding rectangle."""
    h, w = prob.shape[:2]
    b = np.array(box, dtype=np.float32).copy()
    xmin = int(np.clip(np.floor(b[:, 0].min()), 0, w - 1)); xmax = int(np.clip(np.ceil(b[:, 0].max()), 0, w - 1))
    ymin = int(np.clip(np.floor(b[:, 1].min()), 0, h - 1)); ymax = int(np.clip(np.ceil(b[:, 1].max()), 0, h - 1))
    b[:, 0] -= xmin
    b[:, 1] -= ymin
    mask = fill_poly(ymax - ymin + 1, xmax - xmin + 1, b.astype(np.int32))
    roi = prob[ymin:ymax + 1, xmin:xmax + 1].astype(np.float64)
    cnt = int(mask.sum())
    return float((roi * mask).sum() / cnt) if cnt else 0.0
