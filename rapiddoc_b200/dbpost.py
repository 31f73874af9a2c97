"""Host half of DBPostProcess for a whole WINDOW of pages (product code; no oracle import).

Reference: rapidocr `DBPostProcess` as patched by rapid_doc/model/ocr/ocr_patch.py:223-241 (`__call__`), :161-172
(`unclip`), ctor defaults :145-153; upstream `boxes_from_bitmap` / `get_mini_boxes` / `box_score_fast` /
`filter_det_res` (PaddleOCR db_postprocess.py).  Division of labour on the B200 path:

  GPU    binarise + 2x2 dilate (det engine), box_score_fast over the device-resident prob map (rdb_db_box_scores)
  host   cv2.findContours / cv2.minAreaRect on the 1-byte bitmap (the same OpenCV calls the reference makes, run for all
         pages of the window on a thread pool — OpenCV releases the GIL), Clipper offset in C++ (rdb_clipper_offset),
         and everything that is plain arithmetic on [n,4,2] arrays VECTORISED over all boxes of a page with the
         reference's operation order and dtypes, so results are identical to the per-box loops.

`score_fn(quads[m,4,2] f32, page_idx[m]) -> scores[m] f64` is injected: the detector passes the GPU scorer, the CPU
tests pass cv2's.
"""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor
import os

import cv2
import numpy as np

from . import _lib

_POOL = None
TIMES = {}          # cumulative wall-clock seconds per host stage (diagnostics: tools/pipeline_profile.py prints them)


class timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        import time
        self.t = time.perf_counter()

    def __exit__(self, *a):
        import time
        TIMES[self.name] = TIMES.get(self.name, 0.0) + time.perf_counter() - self.t


def host_threads():
    """Host threads this process may use: its CPU affinity (one process per GPU under torchrun gets its share of the box),
    not the machine's core count."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 8


def pool():
    """Process-wide worker pool for the per-page OpenCV calls."""
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=max(2, min(32, host_threads())))
    return _POOL


def mini_boxes(rect_pts):
    """get_mini_boxes' corner ordering for k boxes at once.  rect_pts [k,4,2] f32 = cv2.boxPoints of each min-area rect.
    Python's sorted(key=x) is stable, so is the argsort here."""
    if len(rect_pts) == 0:
        return rect_pts.reshape(0, 4, 2)
    order = np.argsort(rect_pts[:, :, 0], axis=1, kind="stable")
    p = np.take_along_axis(rect_pts, order[:, :, None], axis=1)
    first = p[:, 1, 1] > p[:, 0, 1]
    i1 = np.where(first, 0, 1)
    i4 = np.where(first, 1, 0)
    second = p[:, 3, 1] > p[:, 2, 1]
    i2 = np.where(second, 2, 3)
    i3 = np.where(second, 3, 2)
    idx = np.stack([i1, i2, i3, i4], axis=1)
    return np.take_along_axis(p, idx[:, :, None], axis=1)


def contour_area_f32(boxes):
    """cv2.contourArea for k float32 quads: double accumulation of (double)prev.x * p.y - (double)prev.y * p.x in vertex order
    (OpenCV shapedescr.cpp contourArea), |0.5 * a|."""
    b = boxes.astype(np.float64)
    prev = np.roll(b, 1, axis=1)
    t = prev[:, :, 0] * b[:, :, 1] - prev[:, :, 1] * b[:, :, 0]
    a = t[:, 0].copy()
    for i in range(1, b.shape[1]):
        a = a + t[:, i]
    return np.abs(a * 0.5)


def arc_length_f32(boxes):
    """cv2.arcLength(box, True) for k float32 quads: float32 dx*dx + dy*dy and sqrt per edge, summed into a double in OpenCV's
    order (the buffer is drained back to front)."""
    b = boxes.astype(np.float32)
    prev = np.roll(b, 1, axis=1)
    dx = b[:, :, 0] - prev[:, :, 0]
    dy = b[:, :, 1] - prev[:, :, 1]
    seg = np.sqrt(dx * dx + dy * dy).astype(np.float64)
    p = np.zeros(len(b), np.float64)
    for i in range(b.shape[1] - 1, -1, -1):
        p = p + seg[:, i]
    return p


def clipper_offset(box, distance, cap=512):
    """DBPostProcess.unclip's pyclipper call (ocr_patch.py:166-171) on the library's native Clipper restatement."""
    xy = (C.c_double * 8)(*[float(v) for v in np.asarray(box, np.float64).reshape(-1)])
    out = (C.c_int64 * (2 * cap))()
    n = _lib.check(_lib.load().rdb_clipper_offset(xy, 4, float(distance), out, cap))
    return np.frombuffer(out, dtype=np.int64)[: 2 * n].reshape(-1, 1, 2).astype(np.int32)


def order_points_clockwise(boxes):
    """filter_det_res' order_points_clockwise for k boxes [k,4,2] (np.argsort on 4 elements is an insertion sort: stable)."""
    o = np.argsort(boxes[:, :, 0], axis=1, kind="stable")
    xs = np.take_along_axis(boxes, o[:, :, None], axis=1)
    left, right = xs[:, :2], xs[:, 2:]
    lo = np.argsort(left[:, :, 1], axis=1, kind="stable")
    left = np.take_along_axis(left, lo[:, :, None], axis=1)
    ro = np.argsort(right[:, :, 1], axis=1, kind="stable")
    right = np.take_along_axis(right, ro[:, :, None], axis=1)
    return np.stack([left[:, 0], right[:, 0], right[:, 1], left[:, 1]], axis=1).astype(np.float32)


def find_contours_window(bitmaps, max_threads=0):
    """cv2.findContours(bm, RETR_LIST, CHAIN_APPROX_SIMPLE) for every page of a window in one native call (host threads, no GIL;
    rdb_contours_*: same contours, order and points as OpenCV).  bitmaps [n,h,w] uint8 C-contiguous.
    Returns (per_page counts [n], sizes [total], points [total_points, 1, 2] int32)."""
    lib = _lib.load()
    bitmaps = np.ascontiguousarray(bitmaps, np.uint8)
    n, h, w = bitmaps.shape
    hnd = C.c_void_p()
    _lib.check(lib.rdb_contours_trace(_lib.ptr(bitmaps), n, h, w, int(max_threads) if max_threads > 0 else host_threads(), C.byref(hnd)))
    try:
        per_page = np.zeros(n, np.int32)
        tc, tp = C.c_int64(), C.c_int64()
        _lib.check(lib.rdb_contours_counts(hnd, _lib.ptr(per_page), C.byref(tc), C.byref(tp)))
        sizes = np.zeros(max(1, tc.value), np.int32)
        pts = np.zeros((max(1, tp.value), 1, 2), np.int32)
        _lib.check(lib.rdb_contours_fetch(hnd, _lib.ptr(sizes), _lib.ptr(pts)))
    finally:
        lib.rdb_contours_free(hnd)
    return per_page, sizes[:tc.value], pts[:tp.value]


def clipper_offset_batch(boxes, distances):
    """rdb_clipper_offset for every box of a window in one C call.  Returns (points int32 [total,1,2], starts int64 [m+1])."""
    m = len(boxes)
    xy = np.ascontiguousarray(boxes, np.float64).reshape(m, 8)
    dist = np.ascontiguousarray(distances, np.float64)
    # JT_ROUND emits ~ pi * sqrt(delta / 0.25) / 2 points per corner; sized generously, regrown on overflow
    cap = int(m * 64 + 4 * float(np.sum(np.sqrt(np.maximum(dist, 0.0)))) * 4) + 64
    counts = np.zeros(m, np.int32)
    lib = _lib.load()
    while True:
        out = np.empty((cap, 2), np.int64)
        rc = lib.rdb_clipper_offset_batch(_lib.ptr(xy), m, _lib.ptr(dist), _lib.ptr(out), cap, _lib.ptr(counts))
        if rc >= 0:
            break
        if cap > (1 << 26):
            _lib.check(rc)
        cap *= 4
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return out[:rc].astype(np.int32).reshape(-1, 1, 2), starts


def window_boxes(bitmaps, ori_shapes, score_fn, box_thresh=0.5, unclip_ratio=1.6, max_candidates=1000, min_size=3,
                 parallel=True):
    """DBPostProcess for every page of a window.  bitmaps [n,h,w] uint8 (host), ori_shapes n x (src_h, src_w).
    Returns [(boxes [k,4,2] f32, scores)] per page, identical to running the reference's per-page loop.

    Contours come from the native border follower (rdb_contours_*, host threads, no GIL, identical to cv2.findContours); the
    per-contour cv2.minAreaRect / cv2.boxPoints calls stay on the calling thread (microseconds each — Python threads would only
    fight over the GIL), and all array arithmetic is done ONCE for the boxes of the whole window."""
    n = len(bitmaps)
    height, width = bitmaps[0].shape
    if len(set(tuple(s) for s in ori_shapes)) != 1:      # mixed source sizes: page by page (the reference's own loop)
        return [window_boxes(bitmaps[i:i + 1], ori_shapes[i:i + 1], _shift_pages(score_fn, i), box_thresh, unclip_ratio, max_candidates,
                             min_size, False)[0] for i in range(n)]
    src_h, src_w = ori_shapes[0]
    with timed("det.findContours(native)"):
        # every page of the window traced by the native border follower on host threads (no GIL): same contours, order and
        # points as cv2.findContours(RETR_LIST, CHAIN_APPROX_SIMPLE) (tests/test_contours.py)
        per_page, sizes, cpts = find_contours_window(bitmaps, 0 if parallel else 1)
    with timed("det.minAreaRect"):
        pts, ss, page_idx = [], [], []
        so = po = 0
        for i in range(n):
            k = int(per_page[i])
            for j in range(k):
                m = int(sizes[so + j])
                if j < max_candidates:
                    r = cv2.minAreaRect(cpts[po: po + m])
                    ss.append(min(r[1]))
                    pts.append(cv2.boxPoints(r))
                    page_idx.append(i)
                po += m
            so += k
    empty = (np.zeros((0, 4, 2), np.float32), [])
    if not pts:
        return [empty for _ in range(n)]
    ss = np.array(ss)
    keep = np.nonzero(~(ss < min_size))[0]
    if len(keep) == 0:
        return [empty for _ in range(n)]
    quads = np.ascontiguousarray(mini_boxes(np.stack(pts)[keep]), np.float32)
    page_idx = np.asarray(page_idx, np.int32)[keep]
    with timed("det.box_scores(gpu)"):
        scores = np.asarray(score_fn(quads, page_idx), np.float64)
    with timed("det.unclip+filter"):
        keep = np.nonzero(~(box_thresh > scores))[0]
        quads, scores, page_idx = quads[keep], scores[keep], page_idx[keep]
        area = contour_area_f32(quads)
        length = arc_length_f32(quads)
        ok = length > 0
        quads, scores, page_idx, area, length = quads[ok], scores[ok], page_idx[ok], area[ok], length[ok]
        if len(quads) == 0:
            return [empty for _ in range(n)]
        poly, starts = clipper_offset_batch(quads, area * unclip_ratio / length)
        out_pts, sel = [], []
        for i in range(len(quads)):
            if starts[i + 1] == starts[i]:
                continue
            rect = cv2.minAreaRect(poly[starts[i]:starts[i + 1]])
            if min(rect[1]) < min_size + 2:
                continue
            out_pts.append(cv2.boxPoints(rect))
            sel.append(i)
        if not out_pts:
            return [empty for _ in range(n)]
        sel = np.asarray(sel)
        scores, page_idx = scores[sel], page_idx[sel]
        box = mini_boxes(np.stack(out_pts))                      # float32
        box[:, :, 0] = np.clip(np.round(box[:, :, 0] / width * src_w), 0, src_w)
        box[:, :, 1] = np.clip(np.round(box[:, :, 1] / height * src_h), 0, src_h)
        box = box.astype(np.int32)
        # filter_det_res
        b = order_points_clockwise(box)
        b[:, :, 0] = np.clip(b[:, :, 0], 0, src_w - 1).astype(np.int64)
        b[:, :, 1] = np.clip(b[:, :, 1], 0, src_h - 1).astype(np.int64)
        d01 = b[:, 0] - b[:, 1]
        d03 = b[:, 0] - b[:, 3]
        rw = np.sqrt((d01 * d01).sum(1)).astype(np.int64)        # int(np.linalg.norm(...)) on float32 vectors
        rh = np.sqrt((d03 * d03).sum(1)).astype(np.int64)
        ok = ~((rw <= 3) | (rh <= 3))
        b, scores, page_idx = b[ok], scores[ok], page_idx[ok]
        cuts = np.searchsorted(page_idx, np.arange(n + 1))           # page_idx is non-decreasing (contours were visited page by page)
        return [(b[cuts[i]:cuts[i + 1]], [float(v) for v in scores[cuts[i]:cuts[i + 1]]]) for i in range(n)]


def _shift_pages(score_fn, page):
    return lambda quads, page_idx: score_fn(quads, page_idx + page)


def cv2_score_fn(prob):
    """box_score_fast with cv2 itself over host prob maps [n,h,w] — the scorer of the CPU tests and the fallback for quads the
    GPU scorer flags (vertex outside its clipped bounding box)."""
    def score_one(p, box):
        h, w = p.shape[:2]
        b = box.copy()
        xmin = int(np.clip(np.floor(b[:, 0].min()), 0, w - 1)); xmax = int(np.clip(np.ceil(b[:, 0].max()), 0, w - 1))
        ymin = int(np.clip(np.floor(b[:, 1].min()), 0, h - 1)); ymax = int(np.clip(np.ceil(b[:, 1].max()), 0, h - 1))
        mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
        b[:, 0] -= xmin
        b[:, 1] -= ymin
        cv2.fillPoly(mask, b.reshape(1, -1, 2).astype(np.int32), 1)
        return cv2.mean(p[ymin:ymax + 1, xmin:xmax + 1], mask)[0]

    def fn(quads, page_idx):
        return np.array([score_one(prob[int(pi)], q) for q, pi in zip(quads, page_idx)], np.float64)
    fn.score_one = score_one
    return fn


def gpu_score_fn(device, prob_dev, n, h, w, stream=None):
    """rdb_db_box_scores over the device-resident prob maps (torch tensor or raw device pointer); quads whose vertices leave
    their clipped bounding box are scored with cv2 on the ROI fetched from the device.  stream: raw cudaStream_t of a stream
    that does not wait on the detector's next sub-batch (the legacy default stream would)."""
    lib = _lib.load()

    def fn(quads, page_idx):
        m = len(quads)
        scores = np.zeros(m, np.float64)
        flags = np.zeros(m, np.int32)
        with timed("det.box_scores.ccall"):
            rc = lib.rdb_db_box_scores(int(device), _lib.ptr(prob_dev), int(n), int(h), int(w), m, _lib.ptr(quads), _lib.ptr(page_idx),
                                       _lib.ptr(scores), _lib.ptr(flags), stream)
        _lib.check(rc)
        bad = np.nonzero(flags)[0]
        if len(bad):
            one = cv2_score_fn(None).score_one
            for i in bad:
                q = quads[i]
                x0 = int(np.clip(np.floor(q[:, 0].min()), 0, w - 1)); x1 = int(np.clip(np.ceil(q[:, 0].max()), 0, w - 1))
                y0 = int(np.clip(np.floor(q[:, 1].min()), 0, h - 1)); y1 = int(np.clip(np.ceil(q[:, 1].max()), 0, h - 1))
                roi = prob_dev[int(page_idx[i]), y0:y1 + 1, x0:x1 + 1].cpu().numpy()
                # same arithmetic as box_score_fast on the full map: the ROI is the bounding rectangle it would slice out
                full = _RoiView(roi, x0, y0, w, h)
                scores[i] = one(full, q)
        return scores
    return fn


class _RoiView:
    """Looks like the [h,w] prob map to box_score_fast but only holds the bounding rectangle that function reads."""

    def __init__(self, roi, x0, y0, w, h):
        self.roi, self.x0, self.y0, self.shape = roi, x0, y0, (h, w)

    def __getitem__(self, key):
        ys, xs = key
        return self.roi[ys.start - self.y0: ys.stop - self.y0, xs.start - self.x0: xs.stop - self.x0]
