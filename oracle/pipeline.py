"""ORACLE (test infrastructure, never on the product path): the reference's det -> crop -> rec window flow on CPU.

Restates, for a list of BGR pages, what RapidDoc runs between layout and markdown for OCR text:
  `_run_ocr_det_batch`           rapid_doc/backend/pipeline/analyze_utils.py:105-212  (det per page, sorted_boxes,
                                 merge_det_boxes, get_rotate_crop_image per box)
  `_run_ocr_rec_postprocess`     analyze_utils.py:216-292  (ALL crops of the window in one `ocr(det=False)` call)
  `RapidOcrModel.text_recognizer_call`  rapid_doc/model/ocr/rapid_ocr.py:404-472 (sort by w/h, batches of rec_batch_num,
                                 resize_norm_img, session, CTCLabelDecode)
with the networks of oracle/nets.py (bit-identical to the reference's torch modules, tests/test_oracle.py) and the
rapidocr pieces of oracle/ocr_post.py (parity unpinned, see its header).  merge_det_boxes / update_det_boxes come from
rapiddoc_b200/lines.py, which is host-only arithmetic pinned against the reference's own functions by tests/test_lines.py.

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference / parity legs only.
"""
import copy

import cv2
import numpy as np

from oracle import nets, ocr_post as P
from rapiddoc_b200.lines import merge_det_boxes, sorted_boxes, update_det_boxes


def get_rotate_crop_image(img, points):
    """rapid_doc/utils/ocr_utils.py:494-537."""
    points = np.asarray(points, dtype=np.float32)
    cw = int(max(np.linalg.norm(points[0] - points[1]), np.linalg.norm(points[2] - points[3])))
    ch = int(max(np.linalg.norm(points[0] - points[3]), np.linalg.norm(points[1] - points[2])))
    std = np.float32([[0, 0], [cw, 0], [cw, ch], [0, ch]])
    M = cv2.getPerspectiveTransform(points, std)
    dst = cv2.warpPerspective(img, M, (cw, ch), borderMode=cv2.BORDER_REPLICATE, flags=cv2.INTER_CUBIC)
    if dst.shape[0] * 1.0 / dst.shape[1] >= 2:
        dst = np.rot90(dst)
    return dst


def det_pages(pages, limit_side_len=960, box_thresh=0.3, unclip_ratio=1.8, merge=True, mfd_res_list=None, det_batch=4,
              return_maps=False):
    """Per page: DetPreProcess -> DBNet -> DBPostProcess -> sorted_boxes (detector) -> sorted_boxes / merge / update (caller).
    Returns [list of [4,2] boxes] (and the prob maps / bitmaps when asked)."""
    out, maps = [], []
    for b0 in range(0, len(pages), det_batch):
        xs = [P.det_preprocess(p, limit_side_len=limit_side_len) for p in pages[b0:b0 + det_batch]]
        same = all(x is not None and x.shape == xs[0].shape for x in xs)
        probs = nets.det_forward(np.concatenate(xs)) if same else np.concatenate([nets.det_forward(x) for x in xs])
        for k, page in enumerate(pages[b0:b0 + det_batch]):
            prob = probs[k:k + 1]
            boxes, _ = P.db_postprocess(prob, page.shape[:2], 0.3, box_thresh, unclip_ratio, True)
            if return_maps:
                maps.append((prob[0, 0], P.db_bitmap(prob[0, 0], 0.3, True)))
            if len(boxes) == 0:
                out.append([])
                continue
            bl = sorted_boxes(np.array(sorted_boxes(boxes)))
            if merge:
                bl = merge_det_boxes(bl)
            mfd = mfd_res_list[b0 + k] if mfd_res_list else None
            if mfd:
                bl = update_det_boxes(bl, mfd)
            out.append(bl)
    return (out, maps) if return_maps else out


def rec_crops(crops, rec_batch_num=6, return_ids=False):
    """text_recognizer_call: [(text, conf)] in input order (and per-crop argmax ids when asked)."""
    chars = nets.load_characters()
    ratios = [c.shape[1] / float(c.shape[0]) for c in crops]
    order = np.argsort(np.array(ratios))
    res = [("", 0.0)] * len(crops)
    ids = [None] * len(crops)
    for b0 in range(0, len(crops), rec_batch_num):
        idx = order[b0:b0 + rec_batch_num]
        x, _ = P.rec_batch_tensor([crops[i] for i in idx])
        probs = nets.rec_forward(x)
        dec = P.ctc_decode(probs, chars)
        am = probs.argmax(2)
        for j, i in enumerate(idx):
            res[i] = dec[j]
            ids[i] = am[j]
    return (res, ids) if return_ids else res


def ocr_pages(pages, limit_side_len=960, box_thresh=0.3, unclip_ratio=1.8, merge=True, rec_batch_num=6, drop_score=0.5,
              return_detail=False):
    """The window flow: per page None or [[box, (text, score)], ...] with score >= drop_score.
    return_detail: also a dict with the intermediate results (boxes per page, crops, prob maps / bitmaps, per-crop
    (text, conf) and argmax ids) for parity reports."""
    boxes, maps = det_pages(pages, limit_side_len, box_thresh, unclip_ratio, merge, return_maps=True)
    crops, owner = [], []
    for i, (page, bl) in enumerate(zip(pages, boxes)):
        for b in bl:
            crops.append(get_rotate_crop_image(page, copy.deepcopy(np.asarray(b, np.float32))))
            owner.append(i)
    rec, ids = rec_crops(crops, rec_batch_num, return_ids=True) if crops else ([], [])
    out = [[] for _ in pages]
    q = 0
    for i, bl in enumerate(boxes):
        for b in bl:
            t, s = rec[q]
            q += 1
            if s >= drop_score:
                out[i].append([np.asarray(b).tolist(), (t, s)])
    out = [o or None for o in out]
    if return_detail:
        return out, dict(boxes=boxes, crops=crops, maps=maps, rec=rec, ids=ids)
    return out
