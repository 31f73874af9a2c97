"""Text-line box glue between detection and recognition, restated so that the B200 path does not depend on an
importable `rapid_doc` package: `sorted_boxes`, `merge_det_boxes`, `update_det_boxes`.

Reference: rapid_doc/utils/ocr_utils.py:105-127 (sorted_boxes), :257-317 (merge_det_boxes, with merge_spans_to_line :16-38,
_is_overlaps_y_exceeds_threshold :40-52, merge_overlapping_spans :219-254, calculate_is_angle :478-485) and :192-217
(update_det_boxes, with remove_intervals :156-189, merge_intervals :139-153).  tests/test_lines.py runs these against the
reference's own functions (imported from /root/reference when it is mounted) and against committed fixtures.
"""
import numpy as np

LINE_WIDTH_TO_HEIGHT_RATIO_THRESHOLD = 4     # ocr_utils.py:13


def sorted_boxes(dt_boxes):
    """Top-to-bottom, left-to-right with the 10-px row bubble rule."""
    boxes = sorted(dt_boxes, key=lambda b: (b[0][1], b[0][0]))
    for i in range(len(boxes) - 1):
        for j in range(i, -1, -1):
            if abs(boxes[j + 1][0][1] - boxes[j][0][1]) < 10 and boxes[j + 1][0][0] < boxes[j][0][0]:
                boxes[j], boxes[j + 1] = boxes[j + 1], boxes[j]
            else:
                break
    return boxes


def _is_angle(poly):
    p1, p2, p3, p4 = poly
    height = ((p4[1] - p1[1]) + (p3[1] - p2[1])) / 2
    return not (0.8 * height <= (p3[1] - p1[1]) <= 1.2 * height)


def _bbox(points):
    return [points[0][0], points[0][1], points[1][0], points[2][1]]


def _points(bbox):
    x0, y0, x1, y1 = bbox
    return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]]).astype("float32")


def _y_overlap_exceeds(b1, b2, thr):
    overlap = max(0, min(b1[3], b2[3]) - max(b1[1], b2[1]))
    min_h = min(b1[3] - b1[1], b2[3] - b2[1])
    return (overlap / min_h) > thr if min_h > 0 else False


def merge_det_boxes(dt_boxes):
    """Boxes on one text line (y overlap > 0.6 of the lower box, in y0 order) are merged where they overlap in x, but only
    when the line is wider than 4x its height; rotated boxes pass through at the end."""
    flat, angled = [], []
    for box in dt_boxes:
        (angled if _is_angle(box) else flat).append(box)
    spans = [_bbox(b) for b in flat]
    spans.sort(key=lambda s: s[1])                       # stable, like list.sort in merge_spans_to_line
    lines, cur = [], []
    for s in spans:
        if cur and not _y_overlap_exceeds(s, cur[-1], 0.6):
            lines.append(cur)
            cur = []
        cur.append(s)
    if cur:
        lines.append(cur)
    out = []
    for line in lines:
        width = max(s[2] for s in line) - min(s[0] for s in line)
        height = max(s[3] for s in line) - min(s[1] for s in line)
        if width > height * LINE_WIDTH_TO_HEIGHT_RATIO_THRESHOLD:
            line.sort(key=lambda s: s[0])
            merged = []
            for s in line:
                if not merged or merged[-1][2] < s[0]:
                    merged.append(s)
                else:
                    m = merged.pop()
                    merged.append((min(m[0], s[0]), min(m[1], s[1]), max(m[2], s[2]), max(m[3], s[3])))
            out.extend(_points(s) for s in merged)
        else:
            out.extend(_points(s) for s in line)
    out.extend(angled)
    return out


def _remove_intervals(original, masks):
    masks = sorted(masks, key=lambda m: m[0])
    merged = []
    for m in masks:
        if not merged or merged[-1][1] < m[0]:
            merged.append(list(m))
        else:
            merged[-1][1] = max(merged[-1][1], m[1])
    start, end = original
    out = []
    for ms, me in merged:
        if ms > end or me < start:
            continue
        if start < ms:
            out.append([start, ms - 1])
        start = max(me + 1, start)
    if start <= end:
        out.append([start, end])
    return out


def update_det_boxes(dt_boxes, mfd_res):
    """Cut the x-ranges of formula boxes (y overlap > 0.8) out of every upright text box."""
    out, angled = [], []
    for box in dt_boxes:
        if _is_angle(box):
            angled.append(box)
            continue
        tb = _bbox(box)
        masks = [[mf["bbox"][0], mf["bbox"][2]] for mf in mfd_res if _y_overlap_exceeds(tb, mf["bbox"], 0.8)]
        for lo, hi in _remove_intervals([tb[0], tb[2]], masks):
            out.append(_points([lo, tb[1], hi, tb[3]]))
    out.extend(angled)
    return out


def sort_merge_window(boxes_per_page, merge=True):
    """sorted_boxes (detector) -> sorted_boxes (caller) -> merge_det_boxes for every page of a window in one native call
    (rdb_lines_sort_merge: the logic above in float32, off the interpreter).  boxes_per_page: list of [k,4,2] arrays (or empty).
    Returns a list of [k',4,2] float32 arrays."""
    from . import _lib
    counts = [len(b) for b in boxes_per_page]
    offs = np.zeros(len(counts) + 1, np.int32)
    offs[1:] = np.cumsum(counts)
    total = int(offs[-1])
    if total == 0:
        return [np.zeros((0, 4, 2), np.float32) for _ in counts]
    flat = np.ascontiguousarray(np.concatenate([np.asarray(b, np.float32).reshape(-1, 8) for b in boxes_per_page if len(b)]), np.float32)
    out = np.empty((total, 8), np.float32)
    out_offs = np.zeros(len(counts) + 1, np.int32)
    _lib.check(_lib.load().rdb_lines_sort_merge(_lib.ptr(flat), _lib.ptr(offs), len(counts), int(bool(merge)), _lib.ptr(out), _lib.ptr(out_offs)))
    return [out[out_offs[p]: out_offs[p + 1]].reshape(-1, 4, 2) for p in range(len(counts))]
