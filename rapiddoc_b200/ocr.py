"""Facade layer: host-side mirror of RapidDoc's OCR model class for the det+rec hot path.

`B200OcrModel` keeps the constructor arguments, methods and attributes of
`rapid_doc/model/ocr/rapid_ocr.py: RapidOcrModel` that the pipeline uses
(`backend/pipeline/analyze_utils.py:105-292`, `backend/pipeline/model_init.py:14-27,96-120`):

    .ocr(img, det=True, rec=True, mfd_res=None, tqdm_enable=False, ...)       rapid_ocr.py:225-299
    .det_batch_predict(img_list, max_batch_size)  -> [(boxes, elapse)]        rapid_ocr.py:474-540
    .text_recognizer_call / .text_recognizer(...) -> txts, scores              rapid_ocr.py:404-472
    .text_detector, .text_recognizer, .rec_batch_num, .drop_score

but runs the B200 engines underneath: pages go to the GPU as uint8 (normalisation fused into
the first conv), the DB binarise+dilate runs on the GPU, recognition returns decoded ids
(never the [B,T,18710] tensor).  Contour extraction / minAreaRect stay on OpenCV exactly as in
the reference (rapidocr DBPostProcess as patched by rapid_doc/model/ocr/ocr_patch.py:223-241);
the Clipper offset is the library's native restatement (no pyclipper needed).

Nothing here imports oracle/: this is product code.
"""
import copy
import ctypes as C
import math
import time

import cv2
import numpy as np

from . import _lib, weights as W
from .engine import DET_MEAN, DET_STD, DetEngine, RecEngine


# ----------------------------------------------------------------------------- small utils
def sorted_boxes(dt_boxes):
    """rapid_doc/utils/ocr_utils.py:105-127."""
    n = len(dt_boxes)
    boxes = sorted(dt_boxes, key=lambda b: (b[0][1], b[0][0]))
    boxes = list(boxes)
    for i in range(n - 1):
        for j in range(i, -1, -1):
            if abs(boxes[j + 1][0][1] - boxes[j][0][1]) < 10 and boxes[j + 1][0][0] < boxes[j][0][0]:
                boxes[j], boxes[j + 1] = boxes[j + 1], boxes[j]
            else:
                break
    return boxes


def get_rotate_crop_image(img, points):
    """rapid_doc/utils/ocr_utils.py:494-537 (perspective warp, INTER_CUBIC, BORDER_REPLICATE)."""
    points = np.asarray(points, dtype=np.float32)
    assert len(points) == 4
    cw = int(max(np.linalg.norm(points[0] - points[1]), np.linalg.norm(points[2] - points[3])))
    ch = int(max(np.linalg.norm(points[0] - points[3]), np.linalg.norm(points[1] - points[2])))
    std = np.float32([[0, 0], [cw, 0], [cw, ch], [0, ch]])
    M = cv2.getPerspectiveTransform(points, std)
    dst = cv2.warpPerspective(img, M, (cw, ch), borderMode=cv2.BORDER_REPLICATE, flags=cv2.INTER_CUBIC)
    if dst.shape[0] * 1.0 / max(dst.shape[1], 1) >= 2:
        dst = np.rot90(dst)
    return dst


def crop_geometry(points):
    """Host geometry of get_rotate_crop_image: crop size, dst->src homography (as cv::warpPerspective holds it) and the
    rot90 decision — the tiny per-quad 3x3 solves stay on OpenCV, the pixel work goes to rdb_warp_crops."""
    points = np.asarray(points, dtype=np.float32)
    assert len(points) == 4
    cw = int(max(np.linalg.norm(points[0] - points[1]), np.linalg.norm(points[2] - points[3])))
    ch = int(max(np.linalg.norm(points[0] - points[3]), np.linalg.norm(points[1] - points[2])))
    if cw < 1 or ch < 1:
        return None
    std = np.float32([[0, 0], [cw, 0], [cw, ch], [0, ch]])
    minv = cv2.invert(cv2.getPerspectiveTransform(points, std))[1]
    return cw, ch, np.ascontiguousarray(minv, np.float64), int(ch * 1.0 / cw >= 2)


class DeviceCrops:
    """The crops of one page kept in GPU memory: a packed uint8 buffer (torch tensor) + per-crop stored (h, w) and byte offsets.
    The recogniser resizes / packs them on the device (rdb_resize_pack_u8); `numpy(i)` fetches one crop for host-side users."""

    def __init__(self, buf, offsets, shapes, device):
        self.buf, self.offsets, self.shapes, self.device = buf, np.asarray(offsets, np.int64), list(shapes), int(device)

    def __len__(self):
        return len(self.shapes)

    def numpy(self, i):
        h, w = self.shapes[i]
        o = int(self.offsets[i])
        return self.buf[o: o + h * w * 3].cpu().numpy().reshape(h, w, 3)

    def to_list(self):
        host = self.buf.cpu().numpy()
        return [host[int(o): int(o) + h * w * 3].reshape(h, w, 3) for o, (h, w) in zip(self.offsets, self.shapes)]


def get_rotate_crop_images_gpu(img, boxes, device=0, keep_on_device=False):
    """All crops of one page in ONE GPU call (rdb_warp_crops), bit-identical to get_rotate_crop_image per box.
    img [H,W,3] uint8 numpy (or a device tensor); returns a list of numpy crops (None where the quad is degenerate), or — with
    keep_on_device and no degenerate quad — a DeviceCrops whose pixels never visit the host."""
    geo = [crop_geometry(b) for b in boxes]
    live = [g for g in geo if g is not None]
    if not live:
        return [None] * len(geo)
    if keep_on_device and len(live) == len(geo):
        import torch
        lib = _lib.load()
        minv = np.stack([g[2].reshape(9) for g in live])
        sizes = np.array([[g[0], g[1]] for g in live], np.int32)
        rot = np.array([g[3] for g in live], np.int32)
        nbytes = sizes[:, 0].astype(np.int64) * sizes[:, 1] * 3
        offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
        dev = torch.device("cuda", int(device))
        out = torch.empty(int(nbytes.sum()), dtype=torch.uint8, device=dev)
        if isinstance(img, np.ndarray):
            img = torch.from_numpy(np.ascontiguousarray(img, dtype=np.uint8)).to(dev)
        _lib.check(lib.rdb_warp_crops(int(device), _lib.ptr(img), int(img.shape[0]), int(img.shape[1]), len(live), _lib.ptr(minv), _lib.ptr(sizes),
                                      _lib.ptr(rot), _lib.ptr(out), _lib.ptr(offs), int(out.numel()), None))
        shapes = [((g[0], g[1]) if g[3] else (g[1], g[0])) for g in live]      # stored (h, w): rot90 swaps them
        return DeviceCrops(out, offs, shapes, device)
    lib = _lib.load()
    minv = np.stack([g[2].reshape(9) for g in live])
    sizes = np.array([[g[0], g[1]] for g in live], np.int32)
    rot = np.array([g[3] for g in live], np.int32)
    nbytes = sizes[:, 0].astype(np.int64) * sizes[:, 1] * 3
    offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
    out = np.empty(int(nbytes.sum()), np.uint8)
    if isinstance(img, np.ndarray):
        img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = int(img.shape[0]), int(img.shape[1])
    _lib.check(lib.rdb_warp_crops(int(device), _lib.ptr(img), h, w, len(live), _lib.ptr(minv), _lib.ptr(sizes), _lib.ptr(rot),
                                  _lib.ptr(out), _lib.ptr(offs), int(out.size), None))
    crops, k = [], 0
    for g in geo:
        if g is None:
            crops.append(None)
            continue
        cw, ch, _, r = g
        a = out[offs[k]: offs[k] + nbytes[k]]
        crops.append(a.reshape((cw, ch, 3) if r else (ch, cw, 3)))
        k += 1
    return crops


def _reference_line_utils():
    """merge_det_boxes / update_det_boxes are CPU glue of the caller ('negligible');
    when RapidDoc is importable the reference's own functions are used unchanged."""
    try:
        from rapid_doc.utils.ocr_utils import merge_det_boxes, update_det_boxes
        return merge_det_boxes, update_det_boxes
    except Exception:
        return None, None


def unclip_quad(box, unclip_ratio):
    """DBPostProcess.unclip (ocr_patch.py:161-172) on the native Clipper restatement."""
    box = np.asarray(box, dtype=np.float32)
    area = cv2.contourArea(box)
    length = cv2.arcLength(box, True)
    if length <= 0:
        return np.zeros((0, 1, 2), np.int32)
    distance = float(area * unclip_ratio / length)
    xy = (C.c_double * (2 * len(box)))(*box.astype(np.float64).reshape(-1))
    cap = 512
    out = (C.c_int64 * (2 * cap))()
    n = _lib.check(_lib.load().rdb_clipper_offset(xy, len(box), distance, out, cap))
    return np.frombuffer(out, dtype=np.int64)[: 2 * n].reshape(-1, 1, 2).astype(np.int32)


# ----------------------------------------------------------------------------- detection
class DBPostProcess:
    """Host half of rapidocr's DBPostProcess (quad mode); binarise + dilate already happened on
    the GPU, so __call__ takes the prob map AND the bitmap.  ctor defaults: ocr_patch.py:145-153."""

    def __init__(self, thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.6, use_dilation=True, score_mode="fast", **_):
        self.thresh, self.box_thresh, self.max_candidates = thresh, box_thresh, max_candidates
        self.unclip_ratio, self.use_dilation, self.min_size = unclip_ratio, use_dilation, 3

    @staticmethod
    def get_mini_boxes(contour):
        rect = cv2.minAreaRect(contour)
        pts = sorted(list(cv2.boxPoints(rect)), key=lambda p: p[0])
        a, d = (0, 1) if pts[1][1] > pts[0][1] else (1, 0)
        b, c = (2, 3) if pts[3][1] > pts[2][1] else (3, 2)
        return np.array([pts[a], pts[b], pts[c], pts[d]]), min(rect[1])

    @staticmethod
    def box_score_fast(prob, box):
        h, w = prob.shape[:2]
        b = box.copy()
        xmin = int(np.clip(np.floor(b[:, 0].min()), 0, w - 1)); xmax = int(np.clip(np.ceil(b[:, 0].max()), 0, w - 1))
        ymin = int(np.clip(np.floor(b[:, 1].min()), 0, h - 1)); ymax = int(np.clip(np.ceil(b[:, 1].max()), 0, h - 1))
        mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
        b[:, 0] -= xmin
        b[:, 1] -= ymin
        cv2.fillPoly(mask, b.reshape(1, -1, 2).astype(np.int32), 1)
        return cv2.mean(prob[ymin:ymax + 1, xmin:xmax + 1], mask)[0]

    def boxes_from_bitmap(self, prob, bitmap, dest_w, dest_h):
        height, width = bitmap.shape
        res = cv2.findContours(bitmap * 255 if bitmap.max() <= 1 else bitmap, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        contours = res[0] if len(res) == 2 else res[1]
        boxes, scores = [], []
        for contour in contours[: self.max_candidates]:
            pts, sside = self.get_mini_boxes(contour)
            if sside < self.min_size:
                continue
            score = self.box_score_fast(prob, pts.reshape(-1, 2))
            if self.box_thresh > score:
                continue
            exp = unclip_quad(pts, self.unclip_ratio)
            if len(exp) == 0:
                continue
            box, sside = self.get_mini_boxes(exp)
            if sside < self.min_size + 2:
                continue
            box = np.array(box)
            box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_w), 0, dest_w)
            box[:, 1] = np.clip(np.round(box[:, 1] / height * dest_h), 0, dest_h)
            boxes.append(box.astype(np.int32))
            scores.append(score)
        return boxes, scores

    @staticmethod
    def order_points_clockwise(pts):
        xs = pts[np.argsort(pts[:, 0]), :]
        left, right = xs[:2], xs[2:]
        tl, bl = left[np.argsort(left[:, 1])]
        tr, br = right[np.argsort(right[:, 1])]
        return np.array([tl, tr, br, bl], dtype="float32")

    def filter_det_res(self, boxes, scores, img_h, img_w):
        keep, keep_s = [], []
        for box, s in zip(boxes, scores):
            box = self.order_points_clockwise(np.asarray(box))
            box[:, 0] = np.clip(box[:, 0], 0, img_w - 1).astype(np.int64)
            box[:, 1] = np.clip(box[:, 1], 0, img_h - 1).astype(np.int64)
            if int(np.linalg.norm(box[0] - box[1])) <= 3 or int(np.linalg.norm(box[0] - box[3])) <= 3:
                continue
            keep.append(box)
            keep_s.append(s)
        return np.array(keep), keep_s

    def __call__(self, prob, bitmap, ori_shape):
        src_h, src_w = ori_shape
        boxes, scores = self.boxes_from_bitmap(prob, bitmap, src_w, src_h)
        return self.filter_det_res(boxes, scores, src_h, src_w)


class TextDetOutput:
    def __init__(self, img=None, boxes=None, scores=None, elapse=0.0):
        self.img, self.boxes, self.scores, self.elapse = img, boxes, scores, elapse


class B200TextDetector:
    """Mirror of rapidocr TextDetector as configured by RapidOcrModel (rapid_ocr.py:59-67)."""

    def __init__(self, engine: DetEngine, limit_side_len=960, limit_type="max", mean=DET_MEAN, std=DET_STD, thresh=0.3,
                 box_thresh=0.5, unclip_ratio=1.6, use_dilation=True, max_candidates=1000):
        self.engine = engine
        self.limit_side_len, self.limit_type, self.mean, self.std = limit_side_len, limit_type, mean, std
        self.postprocess_op = DBPostProcess(thresh, box_thresh, max_candidates, unclip_ratio, use_dilation)

    def target_size(self, h, w):
        """DetPreProcess geometry: limit side, round to /32.  Returns (rh, rw) or None."""
        if self.limit_type == "max":
            ratio = float(self.limit_side_len) / max(h, w) if max(h, w) > self.limit_side_len else 1.0
        else:
            ratio = float(self.limit_side_len) / min(h, w) if min(h, w) < self.limit_side_len else 1.0
        rh, rw = int(h * ratio), int(w * ratio)
        rh, rw = int(round(rh / 32) * 32), int(round(rw / 32) * 32)
        if rh <= 0 or rw <= 0:
            return None
        return rh, rw

    def resize(self, img):
        """Host version of the resize (cv2), kept for callers that want the preprocessed uint8 page."""
        t = self.target_size(*img.shape[:2])
        if t is None:
            return None
        return np.ascontiguousarray(img) if t == img.shape[:2] else cv2.resize(img, (t[1], t[0]))

    def detect_batch(self, imgs):
        """Same-size images -> [(boxes, scores)] (rapid_ocr.py:500-540; the reference requires the
        bucket to be same-size too).  Pages are uploaded as they are (uint8); the cv2.resize of
        DetPreProcess, its normalisation, the network and binarise+dilate all run on the GPU."""
        h, w = imgs[0].shape[:2]
        assert all(im.shape[:2] == (h, w) for im in imgs), "det batch must be same-size (as in the reference)"
        t = self.target_size(h, w)
        if t is None:
            return [(None, []) for _ in imgs]
        pages = np.stack([np.ascontiguousarray(im) for im in imgs])
        po = self.postprocess_op
        prob, bitmap = self.engine.infer_u8(pages, thresh=po.thresh, use_dilation=po.use_dilation, mean=self.mean, std=self.std,
                                            resize_to=t)
        out = []
        for i, im in enumerate(imgs):
            boxes, scores = po(prob[i], bitmap[i], im.shape[:2])
            out.append((boxes, scores))
        return out

    def __call__(self, img):
        t0 = time.perf_counter()
        boxes, scores = self.detect_batch([img])[0]
        if boxes is None or len(boxes) == 0:
            return TextDetOutput(img, None, None, time.perf_counter() - t0)
        boxes = np.array(sorted_boxes(boxes))
        return TextDetOutput(img, boxes, scores, time.perf_counter() - t0)


# ----------------------------------------------------------------------------- recognition
class TextRecOutput:
    def __init__(self, imgs, txts, scores, word_results=None, elapse=0.0):
        self.imgs, self.txts, self.scores, self.word_results, self.elapse = imgs, txts, scores, word_results, elapse


class WordInfo:
    """Same fields as rapidocr.ch_ppocr_rec.typings.WordInfo (consumed by CalRecBoxes.cal_ocr_word_box,
    rapid_doc/model/ocr/ocr_patch.py:261-330)."""

    def __init__(self, words=None, word_cols=None, word_types=None, line_txt_len=0.0, confs=None):
        self.words = words or []
        self.word_cols = word_cols or []
        self.word_types = word_types or []
        self.line_txt_len = line_txt_len
        self.confs = confs or []


def _has_chinese_char(text):
    return any("\u4e00" <= ch <= "\u9fff" for ch in text)


def _word_types():
    try:
        from rapidocr.ch_ppocr_rec.typings import WordType
        return WordType.CN, WordType.EN_NUM
    except Exception:
        return "cn", "en&num"


def get_word_info(text, selection):
    """CTCLabelDecode.get_word_info as patched by rapid_doc/model/ocr/ocr_patch.py:333-389: group the decoded
    characters into words (CN vs EN/number runs, spaces kept as their own word) with their CTC columns."""
    CN, EN_NUM = _word_types()
    word_list, word_col_list, state_list = [], [], []
    word_content, word_col_content = [], []
    valid_col = np.where(selection)[0]
    if len(valid_col) <= 0:
        return WordInfo()
    col_width = np.zeros(valid_col.shape)
    col_width[1:] = valid_col[1:] - valid_col[:-1]
    col_width[0] = min(3 if _has_chinese_char(text[0]) else 2, int(valid_col[0]))
    state = None

    def flush():
        nonlocal word_content, word_col_content
        if word_content:
            word_list.append(word_content)
            word_col_list.append(word_col_content)
            state_list.append(state)
            word_content, word_col_content = [], []

    for c_i, char in enumerate(text):
        if char.isspace():
            flush()
            word_list.append([char])
            word_col_list.append([int(valid_col[c_i])])
            state_list.append(EN_NUM)
            state = None
            continue
        c_state = CN if _has_chinese_char(char) else EN_NUM
        if state is None:
            state = c_state
        if state != c_state or col_width[c_i] > 5:
            flush()
            state = c_state
        word_content.append(char)
        word_col_content.append(int(valid_col[c_i]))
    flush()
    return WordInfo(words=word_list, word_cols=word_col_list, word_types=state_list)


class B200TextRecognizer:
    """Mirror of rapidocr TextRecognizer + RapidOcrModel.text_recognizer_call (rapid_ocr.py:404-472)."""

    def __init__(self, engine: RecEngine, characters=None, rec_batch_num=6, rec_image_shape=(3, 48, 320)):
        self.engine = engine
        self.character = characters if characters is not None else W.load_characters()
        self.rec_batch_num = rec_batch_num
        self.rec_image_shape = list(rec_image_shape)

    def _pack(self, crops, max_wh_ratio):
        """resize_norm_img geometry: height 48, width min(imgW, ceil(48*w/h)); uint8, right part is the
        zero pad the GPU applies after normalisation."""
        _, ih, _ = self.rec_image_shape
        iw = int(ih * max_wh_ratio)
        buf = np.zeros((len(crops), ih, iw, 3), np.uint8)
        vw = np.zeros(len(crops), np.int32)
        for i, im in enumerate(crops):
            h, w = im.shape[:2]
            rw = iw if math.ceil(ih * (w / float(h))) > iw else int(math.ceil(ih * (w / float(h))))
            buf[i, :, :rw] = cv2.resize(im, (rw, ih))
            vw[i] = rw
        return buf, vw

    def _pack_device(self, dc, idx, max_wh_ratio):
        """_pack for crops that live on the GPU: the same geometry on the host, the cv2.resize arithmetic in rdb_resize_pack_u8."""
        import torch
        _, ih, _ = self.rec_image_shape
        iw = int(ih * max_wh_ratio)
        sizes = np.array([[dc.shapes[i][1], dc.shapes[i][0]] for i in idx], np.int32)
        vw = np.array([iw if math.ceil(ih * (w / float(h))) > iw else int(math.ceil(ih * (w / float(h)))) for w, h in sizes], np.int32)
        offs = np.ascontiguousarray(dc.offsets[list(idx)], np.int64)
        buf = torch.empty((len(idx), ih, iw, 3), dtype=torch.uint8, device=dc.buf.device)
        _lib.check(_lib.load().rdb_resize_pack_u8(dc.device, _lib.ptr(dc.buf), int(dc.buf.numel()), len(idx), _lib.ptr(offs), _lib.ptr(sizes),
                                                  _lib.ptr(vw), _lib.ptr(buf), ih, iw, None))
        return buf, vw

    def __call__(self, img_list, return_word_box=False):
        if isinstance(img_list, np.ndarray):
            img_list = [img_list]
        t0 = time.perf_counter()
        n = len(img_list)
        on_dev = isinstance(img_list, DeviceCrops)
        ratios = [w / float(h) for h, w in img_list.shapes] if on_dev else [im.shape[1] / float(im.shape[0]) for im in img_list]
        order = np.argsort(np.array(ratios))
        res = [("", 0.0)] * n
        words = [None] * n
        _, ih, iw = self.rec_image_shape
        for b0 in range(0, n, self.rec_batch_num):
            idx = order[b0: b0 + self.rec_batch_num]
            mx = max([iw / ih] + [ratios[i] for i in idx])
            if on_dev:
                buf, vw = self._pack_device(img_list, idx, mx)
                out = self.engine.infer_u8(buf, vw, outs=self.engine._outs(len(idx), self.engine.tokens(buf.shape[2]), vw, False))
            else:
                buf, vw = self._pack([img_list[i] for i in idx], mx)
                out = self.engine.infer_u8(buf, vw)
            T = out["ids"].shape[1]
            for j, i in enumerate(idx):
                ln = int(out["text_len"][j])
                ids = out["text_ids"][j][:ln]
                text = "".join(self.character[k] for k in ids)
                # CTCLabelDecode: float64 mean of the kept float32 max-probs, rounded to 5 decimals
                sel = np.ones(T, bool)
                sel[1:] = out["ids"][j][1:] != out["ids"][j][:-1]
                sel &= out["ids"][j] != 0
                conf = np.array(out["probs"][j][sel]).tolist() or [0]
                res[i] = (text, float(np.mean(conf).round(5)))
                if return_word_box:
                    # rapidocr CTCLabelDecode(return_word_box=True): word grouping + the CTC length scaled to the
                    # crop's share of the padded batch width (wh_ratio / max_wh_ratio)
                    wi = get_word_info(text, sel) if text else WordInfo()
                    wi.line_txt_len = T * ratios[i] / mx
                    wi.confs = conf
                    words[i] = wi
        txts, scores = (list(zip(*res)) if res else ((), ()))
        return TextRecOutput(img_list, tuple(txts), tuple(scores), tuple(words) if return_word_box else None,
                             time.perf_counter() - t0)


# ----------------------------------------------------------------------------- the model class
class B200OcrModel:
    """Drop-in for `RapidOcrModel` (rapid_doc/model/ocr/rapid_ocr.py:43-162) on one B200."""

    def __init__(self, det_db_box_thresh=0.5, lang=None, ocr_config=None, use_dilation=True, det_db_unclip_ratio=1.8,
                 enable_merge_det_boxes=True, is_seal=False, device=0, precision=None):
        if is_seal:
            raise NotImplementedError("seal OCR (PP-OCRv4 seal det) is outside the B200 hot path; use RapidOcrModel(is_seal=True)")
        cfg = dict(ocr_config or {})
        self.drop_score = 0.5
        self.enable_merge_det_boxes = enable_merge_det_boxes
        self.is_seal = False
        prec = _lib.PREC_FP16 if precision is None else precision
        det = DetEngine(device=device, precision=prec, weights_path=cfg.get("Det.model_path"))
        rec = RecEngine(device=device, precision=prec, weights_path=cfg.get("Rec.model_path"))
        self.text_detector = B200TextDetector(
            det, limit_side_len=cfg.get("Det.limit_side_len", 960), limit_type=cfg.get("Det.limit_type", "max"),
            mean=tuple(cfg.get("Det.mean", DET_MEAN)), std=tuple(cfg.get("Det.std", DET_STD)), thresh=cfg.get("Det.thresh", 0.3),
            box_thresh=cfg.get("Det.box_thresh", det_db_box_thresh), unclip_ratio=cfg.get("Det.unclip_ratio", det_db_unclip_ratio),
            use_dilation=cfg.get("Det.use_dilation", use_dilation))
        # rapidocr's default rec_batch_num is 6; the padded width of a batch is set by its widest crop, so the batch
        # size is part of the numerical contract (results can differ between groupings, in the reference too).
        # Raise "Rec.rec_batch_num" in ocr_config for throughput, exactly as with RapidOcrModel.
        self.text_recognizer = B200TextRecognizer(rec, rec_batch_num=cfg.get("Rec.rec_batch_num", 6))
        self.rec_batch_num = self.text_recognizer.rec_batch_num
        self._merge, self._update = _reference_line_utils()

    # ---- rapid_ocr.py:474-540
    def det_batch_predict(self, img_list, max_batch_size=8):
        if not img_list:
            return []
        out = []
        for i in range(0, len(img_list), max_batch_size):
            batch = img_list[i:i + max_batch_size]
            t0 = time.time()
            res = self.text_detector.detect_batch(batch)
            el = (time.time() - t0) / len(batch)
            for boxes, _ in res:
                if boxes is None:
                    out.append((None, 0))
                else:
                    out.append((np.array(sorted_boxes(boxes)) if len(boxes) else boxes, el))
        return out

    def _post_boxes(self, dt_boxes, mfd_res):
        dt_boxes = sorted_boxes(dt_boxes)
        if self.enable_merge_det_boxes and self._merge is not None:
            dt_boxes = self._merge(dt_boxes)
        if mfd_res and self._update is not None:
            dt_boxes = self._update(dt_boxes, mfd_res)
        return dt_boxes

    # ---- rapid_ocr.py:351-401
    def __call__(self, img, mfd_res=None):
        if img is None:
            return None, None
        ori = img.copy()
        det = self.text_detector(img)
        if det.boxes is None:
            return None, None
        dt_boxes = self._post_boxes(det.boxes, mfd_res)
        # opt-in (model.gpu_crop = True): bit-identical results, but at a few dozen lines per page the per-call overheads of the
        # crop entry points still outweigh OpenCV's ~50 us per box (tools/pipeline_probe.py: 20.4 vs 10.9 ms on a 14-line page,
        # 29.0 vs 30.8 ms on a 24-line 1024x1024 page) — it pays once crops are batched over many pages
        if getattr(self, "gpu_crop", False) and len(dt_boxes):
            # all quads of the page in one rdb_warp_crops call (bit-identical to cv2.warpPerspective per box)
            # ... and the crops stay on the device: resize + batch packing happen there too (rdb_resize_pack_u8)
            crops = get_rotate_crop_images_gpu(ori, dt_boxes, self.text_detector.engine.device, keep_on_device=True)
            if not isinstance(crops, DeviceCrops):
                crops = [c if c is not None else get_rotate_crop_image(ori, copy.deepcopy(b)) for c, b in zip(crops, dt_boxes)]
        else:
            crops = [get_rotate_crop_image(ori, copy.deepcopy(b)) for b in dt_boxes]
        rec = self.text_recognizer(crops)
        boxes, res = [], []
        for box, r in zip(dt_boxes, zip(rec.txts, rec.scores)):
            if r[1] >= self.drop_score:
                boxes.append(box)
                res.append(r)
        return boxes, res

    # ---- rapid_ocr.py:225-299
    def ocr(self, img, det=True, rec=True, mfd_res=None, tqdm_enable=False, tqdm_desc="OCR-rec Predict", return_word_box=False,
            ori_img=None, dt_boxes=None):
        assert isinstance(img, (np.ndarray, list))
        if isinstance(img, list) and det:
            raise ValueError("When input a list of images, det must be false")
        if det and rec:
            boxes, res = self.__call__(img, mfd_res=mfd_res)
            if not boxes and not res:
                return [None]
            return [[[np.asarray(b).tolist(), r] for b, r in zip(boxes, res)]]
        if det and not rec:
            d = self.text_detector(img)
            if d.boxes is None:
                return [None]
            boxes = self._post_boxes(np.array(d.boxes), mfd_res)
            return [[np.asarray(b).tolist() for b in boxes]]
        crops = img if isinstance(img, list) else [img]
        r = self.text_recognizer(crops, return_word_box=return_word_box)
        if return_word_box and ori_img is not None and dt_boxes:
            return [list(zip(r.txts, r.scores, self.calc_word_boxes(crops, dt_boxes, r, ori_img.shape[0], ori_img.shape[1])))]
        return [list(zip(r.txts, r.scores))]

    def calc_word_boxes(self, crops, dt_boxes, rec_result, raw_h, raw_w):
        """rapid_ocr.py:301-329: per-word boxes through rapidocr's CalRecBoxes (as patched by ocr_patch.py:261-330).
        CalRecBoxes is rapidocr code that RapidDoc ships with; it is used as-is when importable."""
        try:
            from rapidocr.cal_rec_boxes import CalRecBoxes
        except Exception as exc:  # pragma: no cover
            raise NotImplementedError("return_word_box needs rapidocr's CalRecBoxes (installed with RapidDoc)") from exc
        try:
            from rapid_doc.model.ocr.ocr_patch import patch_word_box
            patch_word_box()
        except Exception:
            pass
        boxes = [np.array(b, dtype=np.float32) for b in dt_boxes]
        out = CalRecBoxes()(crops, boxes, rec_result, False)
        words = []
        for line in out.word_results:
            item = []
            for txt, score, bbox in line:
                if bbox is None:
                    continue
                pts = np.array([bbox]).astype(np.float64)
                pts = np.where(pts < 0, 0, pts)
                pts[..., 0] = np.minimum(pts[..., 0], raw_w)
                pts[..., 1] = np.minimum(pts[..., 1], raw_h)
                item.append((txt, score, pts.astype(np.int32).tolist()[0]))
            if item:
                words.append(tuple(item))
        return tuple(words)

    def text_recognizer_call(self, args, tqdm_enable=False, tqdm_desc="OCR-rec Predict"):
        imgs = [args.img] if isinstance(args.img, np.ndarray) else args.img
        return self.text_recognizer(imgs, getattr(args, "return_word_box", False))
