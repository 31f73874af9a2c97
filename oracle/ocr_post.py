"""ORACLE (test infrastructure, never on the product path): CPU restatement of the OCR
pre/post-processing RapidDoc's hot path performs around the two networks.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this file.

PARITY UNPINNED for the pieces that live in third-party packages absent from
/root/reference:
  * rapidocr (pinned >=3.4.0,<=3.9.0, pyproject.toml:38): DetPreProcess, DBPostProcess
    {boxes_from_bitmap, get_mini_boxes, box_score_fast, unclip, filter_det_res,
    order_points_clockwise, clip_det_res}, TextDetector.sorted_boxes,
    TextRecognizer.resize_norm_img, CTCLabelDecode.  Restated from the published
    PaddleOCR / RapidOCR algorithm as constrained by the in-repo call sites:
    rapid_doc/model/ocr/ocr_patch.py:33-40,145-153,161-172,223-256 and
    rapid_doc/model/ocr/rapid_ocr.py:59-62,404-472,500-540.
  * pyclipper (Clipper 6.4.2, un-pinned transitive): ClipperOffset JT_ROUND /
    ET_CLOSEDPOLYGON restated from the published algorithm (AddPath / FixOrientations /
    DoOffset / OffsetPoint / DoRound, arc_tolerance 0.25, integer rounding half away
    from zero).  The clean-up union Clipper runs afterwards is NOT restated: for the
    convex quads DB produces it does not change the vertex SET, and the only consumer
    (cv2.minAreaRect) depends on the convex hull of that set alone.
The reference holds no golden vectors for any of this (tests/unittest/test_e2e.py only
fuzzy-matches final strings), so these functions are anchored on the call sites above.
cv2 is used exactly where the reference uses cv2.
"""
import math

import cv2
import numpy as np

# ----------------------------------------------------------------------------- det pre


def det_preprocess(img, limit_side_len=960, limit_type="max",
                   mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """DetPreProcess.__call__ (ctor pinned ocr_patch.py:35-37, params
    rapid_ocr.py:59-62).  BGR uint8 HWC -> [1,3,rh,rw] f32 or None."""
    h, w = img.shape[:2]
    if limit_type == "max":
        ratio = float(limit_side_len) / max(h, w) if max(h, w) > limit_side_len else 1.0
    else:
        ratio = float(limit_side_len) / min(h, w) if min(h, w) < limit_side_len else 1.0
    rh, rw = int(h * ratio), int(w * ratio)
    rh = int(round(rh / 32) * 32)
    rw = int(round(rw / 32) * 32)
    if rh <= 0 or rw <= 0:
        return None
    img = cv2.resize(img, (rw, rh))
    scale = np.float32(1.0 / 255.0)
    x = (img.astype(np.float32) * scale - np.array(mean, np.float32)) / np.array(std, np.float32)
    return x.transpose(2, 0, 1)[None].astype(np.float32)


# ----------------------------------------------------------------------------- clipper


def _cround(v):
    """Clipper's Round(): half away from zero, C truncation."""
    return int(v - 0.5) if v < 0 else int(v + 0.5)


def clipper_offset_round(path, delta, arc_tolerance=0.25):
    """ClipperOffset.AddPath(path, JT_ROUND, ET_CLOSEDPOLYGON); Execute(delta) for ONE
    closed polygon, delta > 0.  path: iterable of (x, y); coordinates are truncated to
    integers the way pyclipper's IntPoint conversion does.  Returns int64 [n,2]."""
    pts = [(int(p[0]), int(p[1])) for p in path]
    hi = len(pts) - 1
    if hi < 0:
        return np.zeros((0, 2), np.int64)
    while hi > 0 and pts[0] == pts[hi]:
        hi -= 1
    src = [pts[0]]
    for i in range(1, hi + 1):
        if src[-1] != pts[i]:
            src.append(pts[i])
    if len(src) < 3:
        return np.zeros((0, 2), np.int64)
    # FixOrientations: Orientation() == Area >= 0, else reverse
    a = 0.0
    j = len(src) - 1
    for i in range(len(src)):
        a += (float(src[j][0]) + src[i][0]) * (float(src[j][1]) - src[i][1])
        j = i
    if -a * 0.5 < 0:
        src.reverse()
    n = len(src)
    if abs(delta) < 1e-20:
        return np.array(src, np.int64)
    y = arc_tolerance if arc_tolerance > 0 else 0.25
    if y > abs(delta) * 0.25:
        y = abs(delta) * 0.25
    steps = math.pi / math.acos(1 - y / abs(delta))
    if steps > abs(delta) * math.pi:
        steps = abs(delta) * math.pi
    m_sin = math.sin(2 * math.pi / steps)
    m_cos = math.cos(2 * math.pi / steps)
    steps_per_rad = steps / (2 * math.pi)
    if delta < 0:
        m_sin = -m_sin
    normals = []
    for i in range(n):
        p1, p2 = src[i], src[(i + 1) % n]
        dx, dy = float(p2[0] - p1[0]), float(p2[1] - p1[1])
        f = 1.0 / math.sqrt(dx * dx + dy * dy)
        normals.append((dy * f, -dx * f))
    out = []
    k = n - 1
    for j in range(n):
        nk, nj = normals[k], normals[j]
        sin_a = nk[0] * nj[1] - nj[0] * nk[1]
        if abs(sin_a * delta) < 1.0:
            cos_a = nk[0] * nj[0] + nj[1] * nk[1]
            if cos_a > 0:     # OffsetPoint returns here, before its trailing `k = j`
                out.append((_cround(src[j][0] + nk[0] * delta), _cround(src[j][1] + nk[1] * delta)))
                continue
        elif sin_a > 1.0:
            sin_a = 1.0
        elif sin_a < -1.0:
            sin_a = -1.0
        if True:
            if sin_a * delta < 0:
                out.append((_cround(src[j][0] + nk[0] * delta), _cround(src[j][1] + nk[1] * delta)))
                out.append(src[j])
                out.append((_cround(src[j][0] + nj[0] * delta), _cround(src[j][1] + nj[1] * delta)))
            else:  # DoRound
                ang = math.atan2(sin_a, nk[0] * nj[0] + nk[1] * nj[1])
                st = max(_cround(steps_per_rad * abs(ang)), 1)
                X, Y = nk
                for _ in range(st):
                    out.append((_cround(src[j][0] + X * delta), _cround(src[j][1] + Y * delta)))
                    X2 = X
                    X = X * m_cos - m_sin * Y
                    Y = X2 * m_sin + Y * m_cos
                out.append((_cround(src[j][0] + nj[0] * delta), _cround(src[j][1] + nj[1] * delta)))
        k = j
    return np.array(out, np.int64)


# ----------------------------------------------------------------------------- DB post


def get_mini_boxes(contour):
    bounding_box = cv2.minAreaRect(contour)
    points = sorted(list(cv2.boxPoints(bounding_box)), key=lambda x: x[0])
    if points[1][1] > points[0][1]:
        i1, i4 = 0, 1
    else:
        i1, i4 = 1, 0
    if points[3][1] > points[2][1]:
        i2, i3 = 2, 3
    else:
        i2, i3 = 3, 2
    box = np.array([points[i1], points[i2], points[i3], points[i4]])
    return box, min(bounding_box[1])


def box_score_fast(bitmap, _box):
    h, w = bitmap.shape[:2]
    box = _box.copy()
    xmin = np.clip(np.floor(box[:, 0].min()).astype(np.int32), 0, w - 1)
    xmax = np.clip(np.ceil(box[:, 0].max()).astype(np.int32), 0, w - 1)
    ymin = np.clip(np.floor(box[:, 1].min()).astype(np.int32), 0, h - 1)
    ymax = np.clip(np.ceil(box[:, 1].max()).astype(np.int32), 0, h - 1)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
    box[:, 0] = box[:, 0] - xmin
    box[:, 1] = box[:, 1] - ymin
    cv2.fillPoly(mask, box.reshape(1, -1, 2).astype(np.int32), 1)
    return cv2.mean(bitmap[ymin:ymax + 1, xmin:xmax + 1], mask)[0]


def unclip(box, unclip_ratio):
    """distance = area * ratio / perimeter (ocr_patch.py:161-172), JT_ROUND offset."""
    area = cv2.contourArea(box)
    length = cv2.arcLength(box, True)
    distance = area * unclip_ratio / length
    return clipper_offset_round(box, distance).reshape(-1, 1, 2)


def db_bitmap(pred, thresh=0.3, use_dilation=True):
    """ocr_patch.py:228-235: pred[H,W] f32 -> uint8 {0,1} mask."""
    seg = pred > thresh
    if use_dilation:
        return cv2.dilate(np.array(seg).astype(np.uint8), np.array([[1, 1], [1, 1]]))
    return seg.astype(np.uint8)


def boxes_from_bitmap(pred, bitmap, dest_width, dest_height, box_thresh=0.5, unclip_ratio=1.6,
                      max_candidates=1000, min_size=3):
    height, width = bitmap.shape
    outs = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    contours = outs[0] if len(outs) == 2 else outs[1]
    boxes, scores = [], []
    for contour in contours[:max_candidates]:
        points, sside = get_mini_boxes(contour)
        if sside < min_size:
            continue
        score = box_score_fast(pred, points.reshape(-1, 2))
        if box_thresh > score:
            continue
        exp = unclip(points, unclip_ratio)
        if len(exp) == 0:
            continue
        box, sside = get_mini_boxes(exp.astype(np.int32))
        if sside < min_size + 2:
            continue
        box = np.array(box)
        box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_width), 0, dest_width)
        box[:, 1] = np.clip(np.round(box[:, 1] / height * dest_height), 0, dest_height)
        boxes.append(box.astype(np.int32))
        scores.append(score)
    return np.array(boxes, dtype=np.int32).reshape(-1, 4, 2), scores


def order_points_clockwise(pts):
    xs = pts[np.argsort(pts[:, 0]), :]
    left, right = xs[:2, :], xs[2:, :]
    left = left[np.argsort(left[:, 1]), :]
    tl, bl = left
    right = right[np.argsort(right[:, 1]), :]
    tr, br = right
    return np.array([tl, tr, br, bl], dtype="float32")


def filter_det_res(dt_boxes, scores, img_height, img_width):
    out, out_scores = [], []
    for box, score in zip(dt_boxes, scores):
        box = order_points_clockwise(box)
        for p in range(box.shape[0]):
            box[p, 0] = int(min(max(box[p, 0], 0), img_width - 1))
            box[p, 1] = int(min(max(box[p, 1], 0), img_height - 1))
        rw = int(np.linalg.norm(box[0] - box[1]))
        rh = int(np.linalg.norm(box[0] - box[3]))
        if rw <= 3 or rh <= 3:
            continue
        out.append(box)
        out_scores.append(score)
    return np.array(out), out_scores


def db_postprocess(pred, ori_shape, thresh=0.3, box_thresh=0.5, unclip_ratio=1.6, use_dilation=True):
    """DBPostProcess.__call__ as patched (ocr_patch.py:223-241), quad mode.
    pred: [1,1,H,W] f32 -> (boxes [n,4,2] f32, scores)."""
    src_h, src_w = ori_shape
    p = pred[0, 0]
    mask = db_bitmap(p, thresh, use_dilation)
    boxes, scores = boxes_from_bitmap(p, mask, src_w, src_h, box_thresh, unclip_ratio)
    return filter_det_res(boxes, scores, src_h, src_w)


def sorted_boxes(dt_boxes):
    """rapid_doc/utils/ocr_utils.py:105-127 (same rule as rapidocr TextDetector.sorted_boxes)."""
    n = len(dt_boxes)
    b = sorted(dt_boxes, key=lambda x: (x[0][1], x[0][0]))
    b = list(b)
    for i in range(n - 1):
        for j in range(i, -1, -1):
            if abs(b[j + 1][0][1] - b[j][0][1]) < 10 and b[j + 1][0][0] < b[j][0][0]:
                b[j], b[j + 1] = b[j + 1], b[j]
            else:
                break
    return b


# ----------------------------------------------------------------------------- rec pre/post


def resize_norm_img(img, max_wh_ratio, rec_image_shape=(3, 48, 320)):
    """TextRecognizer.resize_norm_img (rapidocr; called rapid_ocr.py:438)."""
    c, ih, iw = rec_image_shape
    assert c == img.shape[2]
    iw = int(ih * max_wh_ratio)
    h, w = img.shape[:2]
    ratio = w / float(h)
    rw = iw if math.ceil(ih * ratio) > iw else int(math.ceil(ih * ratio))
    r = cv2.resize(img, (rw, ih)).astype("float32")
    r = r.transpose((2, 0, 1)) / 255
    r -= 0.5
    r /= 0.5
    pad = np.zeros((c, ih, iw), dtype=np.float32)
    pad[:, :, 0:rw] = r
    return pad


def rec_batch_tensor(img_list, rec_image_shape=(3, 48, 320)):
    """One batch of text_recognizer_call (rapid_ocr.py:423-440)."""
    c, ih, iw = rec_image_shape
    m = iw / ih
    for im in img_list:
        m = max(m, im.shape[1] * 1.0 / im.shape[0])
    return np.concatenate([resize_norm_img(im, m, rec_image_shape)[None] for im in img_list]).astype(np.float32), m


def ctc_decode(preds, characters):
    """CTCLabelDecode.__call__ (greedy).  preds [B,T,V] softmax probs ->
    [(text, conf)], conf = float64 mean of kept max-probs rounded to 5 decimals."""
    idx = preds.argmax(axis=2)
    prob = preds.max(axis=2)
    return ctc_decode_indices(idx, prob, characters)


def ctc_decode_indices(idx, prob, characters):
    res = []
    for b in range(len(idx)):
        sel = np.ones(len(idx[b]), dtype=bool)
        sel[1:] = idx[b][1:] != idx[b][:-1]
        sel &= idx[b] != 0
        conf = np.array(prob[b][sel]).tolist()
        if len(conf) == 0:
            conf = [0]
        text = "".join(characters[i] for i in idx[b][sel])
        res.append((text, np.mean(conf).round(5).tolist()))
    return res
