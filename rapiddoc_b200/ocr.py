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
import ctypes as C
import math
import time

import cv2
import numpy as np

from . import _lib, dbpost, weights as W
from .engine import DET_MEAN, DET_STD, DetEngine, RecEngine
from .lines import merge_det_boxes, sorted_boxes, update_det_boxes


# ----------------------------------------------------------------------------- small utils
def check_img(img):
    """rapid_doc/utils/ocr_utils.py:73-78: bytes are decoded, 2-D grayscale becomes BGR."""
    if isinstance(img, bytes):
        img = cv2.imdecode(np.frombuffer(img, dtype=np.uint8), cv2.IMREAD_UNCHANGED)
    if isinstance(img, np.ndarray) and img.ndim == 2:
        img = cv2.cvtColor(img, cv2.COLOR_GRAY2BGR)
    return img


def preprocess_image(img):
    """rapid_doc/utils/ocr_utils.py:81-102 (alpha_to_color on white): BGRA is composited onto a white background."""
    if img.ndim == 3 and img.shape[2] == 4:
        B, G, R, A = cv2.split(img)
        alpha = A / 255
        R = (255 * (1 - alpha) + R * alpha).astype(np.uint8)
        G = (255 * (1 - alpha) + G * alpha).astype(np.uint8)
        B = (255 * (1 - alpha) + B * alpha).astype(np.uint8)
        img = cv2.merge((B, G, R))
    return img


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


def crop_geometry_batch(boxes):
    """crop_geometry for all quads of a window.  Returns (keep indices, sizes [k,2] (w,h) i32, minv [k,9] f64, rotate [k] i32).
    Text-line boxes have integer coordinates (DBPostProcess rounds them, merge / update keep integers), for which the four
    float32 np.linalg.norm calls per box reduce to exact integer arithmetic and are evaluated for the whole window at once.
    An axis-aligned box whose crop size equals its extent maps onto the crop by a pure integer translation: cv2's
    getPerspectiveTransform + invert give that matrix up to ~1e-13 noise, which cannot move any 1/32-pixel fixed-point source
    coordinate of warpPerspective (they are integers plus that noise), so the exact translation is used instead of two LU
    solves per box.  Every other quad goes through the same cv2 calls as the reference."""
    m = len(boxes)
    if m == 0:
        return np.zeros(0, np.int64), np.zeros((0, 2), np.int32), np.zeros((0, 9), np.float64), np.zeros(0, np.int32)
    pts = np.asarray(boxes, dtype=np.float32).reshape(m, 4, 2)
    if not np.array_equal(pts, np.round(pts)) or np.abs(pts).max() >= 4096:
        geo = [crop_geometry(b) for b in pts]
        keep = np.array([i for i, g in enumerate(geo) if g is not None], np.int64)
        live = [geo[i] for i in keep]
        return (keep, np.array([[g[0], g[1]] for g in live], np.int32).reshape(-1, 2),
                np.array([g[2].reshape(9) for g in live], np.float64).reshape(-1, 9), np.array([g[3] for g in live], np.int32))

    def norm(a, b):       # float32 sqrt of an exactly representable integer sum of squares == np.linalg.norm(pts[a] - pts[b])
        d = pts[:, a] - pts[:, b]
        return np.sqrt((d * d).sum(1))
    cw = np.maximum(norm(0, 1), norm(2, 3)).astype(np.int64)      # int(max(...)): truncation
    ch = np.maximum(norm(0, 3), norm(1, 2)).astype(np.int64)
    keep = np.nonzero((cw >= 1) & (ch >= 1))[0]
    pts, cw, ch = pts[keep], cw[keep], ch[keep]
    rot = (ch * 1.0 / cw >= 2).astype(np.int32)
    x0, y0, x1, y1 = pts[:, 0, 0], pts[:, 0, 1], pts[:, 2, 0], pts[:, 2, 1]
    aligned = ((pts[:, 1, 1] == y0) & (pts[:, 1, 0] == x1) & (pts[:, 3, 0] == x0) & (pts[:, 3, 1] == y1) &
               (x1 - x0 == cw) & (y1 - y0 == ch))
    minv = np.zeros((len(keep), 9), np.float64)
    minv[:, 0] = 1.0
    minv[:, 4] = 1.0
    minv[:, 8] = 1.0
    minv[:, 2] = x0
    minv[:, 5] = y0
    for i in np.nonzero(~aligned)[0]:
        std = np.float32([[0, 0], [cw[i], 0], [cw[i], ch[i]], [0, ch[i]]])
        minv[i] = cv2.invert(cv2.getPerspectiveTransform(pts[i], std))[1].reshape(9)
    return keep, np.stack([cw, ch], 1).astype(np.int32), minv, rot


class DeviceCrops:
    """The crops of one page kept in GPU memory: a packed uint8 buffer (torch tensor) + per-crop stored (h, w) and byte offsets.
    The recogniser resizes / packs them on the device (rdb_resize_pack_u8); `numpy(i)` fetches one crop for host-side users."""

    def __init__(self, buf, offsets, shapes, device):
        self.buf, self.offsets, self.device = buf, np.asarray(offsets, np.int64), int(device)
        self.shapes = np.asarray(shapes, np.int64).reshape(-1, 2)       # stored (h, w) per crop

    def __len__(self):
        return len(self.shapes)

    def numpy(self, i):
        h, w = (int(v) for v in self.shapes[i])
        o = int(self.offsets[i])
        return self.buf[o: o + h * w * 3].cpu().numpy().reshape(h, w, 3)

    def to_list(self):
        host = self.buf.cpu().numpy()
        return [host[int(o): int(o) + int(h) * int(w) * 3].reshape(int(h), int(w), 3) for o, (h, w) in zip(self.offsets, self.shapes)]


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
        self.stats = {"h2d_bytes": 0, "d2h_bytes": 0, "launches": 0}
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

    def _stage(self, imgs):
        """The same-size pages as ONE contiguous host array [n,h,w,3].  A caller that already holds them in one buffer (e.g.
        views into a pinned window buffer, as a raster producer would fill) pays no copy; otherwise the pages are gathered
        into a grow-only pinned staging buffer on the worker pool (numpy's copy releases the GIL)."""
        import torch
        n = len(imgs)
        h, w = imgs[0].shape[:2]
        if isinstance(imgs, np.ndarray) and imgs.ndim == 4 and imgs.flags.c_contiguous and imgs.dtype == np.uint8:
            return imgs
        step = h * w * 3
        a0 = imgs[0]
        if all(isinstance(im, np.ndarray) and im.dtype == np.uint8 and im.flags.c_contiguous and im.shape == a0.shape and
               im.ctypes.data == a0.ctypes.data + k * step for k, im in enumerate(imgs)):
            return np.lib.stride_tricks.as_strided(a0, shape=(n, h, w, 3), strides=(step,) + a0.strides)   # keeps a0 alive
        need = n * step
        if getattr(self, "_pin", None) is None or self._pin.numel() < need:
            self._pin = torch.empty(need + need // 4, dtype=torch.uint8).pin_memory()
        stage = self._pin.numpy()[:need].reshape(n, h, w, 3)
        list(dbpost.pool().map(lambda k: np.copyto(stage[k], imgs[k]), range(n)))
        return stage

    def detect_window(self, imgs, keep_pages=False, sub_batch=32):
        """Same-size pages -> [(boxes [k,4,2] f32, scores)] (rapid_ocr.py:500-540 for a whole window; the reference requires a
        det batch to be same-size too).  Pages are uploaded as raw uint8 once; DetPreProcess' cv2.resize + normalisation, the
        network and binarise + dilate run on the GPU; the fp32 prob map STAYS on the GPU, where box_score_fast is evaluated
        (rdb_db_box_scores); only the 1-byte bitmap comes back for cv2.findContours.  The window is cut into sub-batches
        so that the host post-processing of one overlaps the GPU work of the next (the GPU call runs on a worker thread;
        ctypes releases the GIL).  keep_pages: also return the device copy of the pages (for the GPU crop path)."""
        import torch
        h, w = int(imgs[0].shape[0]), int(imgs[0].shape[1])
        assert all(tuple(im.shape[:2]) == (h, w) for im in imgs), "det batch must be same-size (as in the reference)"
        n = len(imgs)
        t = self.target_size(h, w)
        if t is None:
            return [(None, []) for _ in imgs], None
        rh, rw = t
        po = self.postprocess_op
        dev = torch.device("cuda", self.engine.device)
        resident = hasattr(imgs, "is_cuda") and imgs.is_cuda           # a [n,h,w,3] uint8 tensor already on this GPU
        if resident:
            assert imgs.dtype == torch.uint8 and imgs.is_contiguous() and imgs.device.index == self.engine.device
            host, pages_dev = None, imgs
        else:
            with dbpost.timed("det.stage"):
                host = self._stage(imgs)
            pages_dev = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
            self.stats["h2d_bytes"] += host.size
        self.stats["d2h_bytes"] += n * rh * rw
        need = n * rh * rw
        if getattr(self, "_pin_bm", None) is None or self._pin_bm.numel() < need:
            self._pin_bm = torch.empty(need + need // 4, dtype=torch.uint8).pin_memory()
        bitmaps = self._pin_bm.numpy()[:need].reshape(n, rh, rw)
        probs = torch.empty((n, rh, rw), dtype=torch.float32, device=dev)

        st_outer = torch.cuda.current_stream(dev)       # the caller's stream: the worker thread below has its own default

        def gpu(lo, hi):
            with torch.cuda.device(dev), torch.cuda.stream(st_outer):
                if host is not None:
                    pages_dev[lo:hi].copy_(torch.from_numpy(host[lo:hi]), non_blocking=True)
                self.engine.infer_u8(pages_dev[lo:hi], thresh=po.thresh, use_dilation=po.use_dilation, mean=self.mean, std=self.std,
                                     resize_to=t, prob=probs[lo:hi], bitmap=bitmaps[lo:hi], stream=torch.cuda.current_stream(dev))
            return lo, hi

        cuts = [(lo, min(n, lo + sub_batch)) for lo in range(0, n, sub_batch)]
        out = [None] * n
        fut = dbpost.pool().submit(gpu, *cuts[0])
        for k, (lo, hi) in enumerate(cuts):
            with dbpost.timed("det.wait_gpu"):
                fut.result()
            self.stats["launches"] += self.engine.last_launches + 1          # + the box scorer
            if k + 1 < len(cuts):
                fut = dbpost.pool().submit(gpu, *cuts[k + 1])
            if getattr(self, "_post_stream", None) is None:
                # high priority: when another window's recogniser keeps the GPU busy (ocr_pages_stream) the tiny scorer
                # kernel must not queue behind its launches
                self._post_stream = torch.cuda.Stream(dev, priority=-1)
            score_fn = dbpost.gpu_score_fn(self.engine.device, probs[lo:hi], hi - lo, rh, rw, self._post_stream.cuda_stream)
            res = dbpost.window_boxes(bitmaps[lo:hi], [(h, w)] * (hi - lo), score_fn, po.box_thresh,
                                      po.unclip_ratio, po.max_candidates, po.min_size)
            out[lo:hi] = res
        return out, (pages_dev if keep_pages else None)

    def detect_batch(self, imgs):
        return self.detect_window(imgs)[0]

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
        self.stats = {"h2d_bytes": 0, "d2h_bytes": 0, "launches": 0, "crops": 0}

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

    def upload(self, crops):
        """Host crops (list of [h,w,3] uint8 arrays of any size) -> DeviceCrops: one packed pinned buffer, one H2D copy."""
        import torch
        shapes = [(int(c.shape[0]), int(c.shape[1])) for c in crops]
        nbytes = np.array([h * w * 3 for h, w in shapes], np.int64)
        offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
        total = int(nbytes.sum())
        if getattr(self, "_pin", None) is None or self._pin.numel() < total:
            self._pin = torch.empty(total + total // 4 + 1024, dtype=torch.uint8).pin_memory()
        stage = self._pin.numpy()

        def put(k):
            h, w = shapes[k]
            np.copyto(stage[offs[k]: offs[k] + nbytes[k]].reshape(h, w, 3), crops[k])
        if len(crops) >= 32:
            list(dbpost.pool().map(put, range(len(crops))))
        else:
            for k in range(len(crops)):
                put(k)
        dev = torch.device("cuda", self.engine.device)
        buf = torch.empty(total, dtype=torch.uint8, device=dev)
        buf.copy_(self._pin[:total], non_blocking=True)
        self.stats["h2d_bytes"] += total
        return DeviceCrops(buf, offs, shapes, self.engine.device)

    def plan(self, shapes):
        """text_recognizer_call's batching (rapid_ocr.py:411-440): crops sorted by w/h, consecutive batches of rec_batch_num,
        every batch padded to int(48 * max(320/48, its widest ratio)).  Returns (order, ratios, [(lo, hi, imgW, max_ratio)])."""
        _, ih, iw = self.rec_image_shape
        sh = np.asarray(shapes, np.float64).reshape(-1, 2)
        ratios = sh[:, 1] / sh[:, 0]                         # w / float(h), the same IEEE division
        order = np.argsort(ratios)
        sorted_r = ratios[order]
        batches = []
        for b0 in range(0, len(sh), self.rec_batch_num):
            b1 = min(len(sh), b0 + self.rec_batch_num)
            mx = max(iw / ih, float(sorted_r[b0:b1].max()))
            batches.append((b0, b1, int(ih * mx), mx))
        return order, ratios, batches

    def __call__(self, img_list, return_word_box=False):
        """All crops of a window in one go (cross-page batching): one upload (unless they already live on the
        GPU), ONE launch that cv2.resizes every crop into its batch slot, then the recogniser once per reference batch —
        enqueued back to back on the stream, nothing synchronises until the (id, prob) pairs of the whole window are
        copied back.  Same batches, same padded widths, same results as the reference's batch loop."""
        import torch
        if isinstance(img_list, np.ndarray):
            img_list = [img_list]
        t0 = time.perf_counter()
        n = len(img_list)
        if n == 0:
            return TextRecOutput(img_list, (), (), () if return_word_box else None, 0.0)
        dc = img_list if isinstance(img_list, DeviceCrops) else self.upload(img_list)
        _, ih, _ = self.rec_image_shape
        order, ratios, batches = self.plan(dc.shapes)
        dev = dc.buf.device
        # slot table in batch order
        sizes = np.ascontiguousarray(np.asarray(dc.shapes, np.int32).reshape(-1, 2)[order][:, ::-1])   # (w, h)
        src_offs = np.ascontiguousarray(dc.offsets[order], np.int64)
        dst_w = np.empty(n, np.int32)
        pitch = np.empty(n, np.int32)
        dst_offs = np.empty(n, np.int64)
        tok_off, base, toks = [], 0, 0
        cw_all = np.ceil(ih * (sizes[:, 0].astype(np.float64) / sizes[:, 1].astype(np.float64)))    # math.ceil(48 * (w / float(h)))
        for lo, hi, imgW, _mx in batches:
            dst_w[lo:hi] = np.minimum(cw_all[lo:hi], imgW).astype(np.int32)
            pitch[lo:hi] = imgW
            dst_offs[lo:hi] = base + np.arange(hi - lo, dtype=np.int64) * (ih * imgW * 3)
            base += (hi - lo) * ih * imgW * 3
            tok_off.append(toks)
            toks += (hi - lo) * self.engine.tokens(imgW)
        with torch.cuda.device(dev):
            packed = torch.empty(base, dtype=torch.uint8, device=dev)
            st = torch.cuda.current_stream(dev)
            _lib.check(_lib.load().rdb_resize_pack_slots(dc.device, _lib.ptr(dc.buf), int(dc.buf.numel()), n, _lib.ptr(src_offs), _lib.ptr(sizes),
                                                         _lib.ptr(dst_w), _lib.ptr(dst_offs), _lib.ptr(pitch), _lib.ptr(packed), int(base), ih,
                                                         st.cuda_stream or None))
            vw_dev = torch.empty(n, dtype=torch.int32, device=dev)
            if getattr(self, "_pin_vw", None) is None or self._pin_vw.numel() < n:
                self._pin_vw = torch.empty(n + 1024, dtype=torch.int32).pin_memory()
            self._pin_vw[:n].copy_(torch.from_numpy(dst_w))
            vw_dev.copy_(self._pin_vw[:n], non_blocking=True)
            ids_all = torch.empty(toks, dtype=torch.int32, device=dev)
            probs_all = torch.empty(toks, dtype=torch.float32, device=dev)
            t_l = dbpost.timed("rec.launch")
            t_l.__enter__()
            for (lo, hi, imgW, _mx), to in zip(batches, tok_off):
                self.engine.infer_u8_raw(packed.data_ptr() + int(dst_offs[lo]), vw_dev.data_ptr() + 4 * lo, hi - lo, imgW,
                                         ids_all.data_ptr() + 4 * to, probs_all.data_ptr() + 4 * to, st.cuda_stream or None)
                self.stats["launches"] += self.engine.last_launches
            t_l.__exit__()
            self.stats["launches"] += 1                                           # resize_slots
            self.stats["h2d_bytes"] += n * (32 + 4)                               # slot table + valid widths
            self.stats["d2h_bytes"] += toks * 8
            self.stats["crops"] += n
            with dbpost.timed("rec.sync+d2h"):
                # pinned destination + stream synchronise: a pageable `.cpu()` is a blocking driver call that also stalls the
                # small copies of the other pipeline stage (see pinned_scratch in csrc/warp.cuh)
                if getattr(self, "_pin_out", None) is None or self._pin_out.numel() < 2 * toks:
                    self._pin_out = torch.empty(2 * toks + 4096, dtype=torch.int32).pin_memory()
                pi, pp = self._pin_out[:toks], self._pin_out[toks: 2 * toks].view(torch.float32)
                pi.copy_(ids_all, non_blocking=True)
                pp.copy_(probs_all, non_blocking=True)
                st.synchronize()
                ids_h = pi.numpy().copy()
                probs_h = pp.numpy().copy()
        if getattr(self, "keep_ids", False):      # diagnostics (bench parity report): per-crop argmax ids in input order
            self.last_ids = [None] * n
            for (lo, hi, imgW, _mx), to in zip(batches, tok_off):
                T = self.engine.tokens(imgW)
                for j in range(hi - lo):
                    self.last_ids[int(order[lo + j])] = ids_h[to + j * T: to + (j + 1) * T].copy()
        with dbpost.timed("rec.decode"):
            res, words = self.decode_window(ids_h, probs_h, batches, tok_off, order, ratios, return_word_box)
        txts, scores = list(zip(*res))
        return TextRecOutput(img_list, tuple(txts), tuple(scores), tuple(words) if return_word_box else None,
                             time.perf_counter() - t0)

    def decode_window(self, ids_h, probs_h, batches, tok_off, order, ratios, return_word_box=False):
        """CTCLabelDecode (rapidocr ch_ppocr_rec/utils.py) for every batch of a window, vectorised per batch: keep t where
        id != previous id and id != blank; text = chars[ids]; conf = float64 mean of the kept float32 max-probs, rounded to
        5 decimals.  ids_h / probs_h: flat host arrays, batch k = [b, T_k] at tok_off[k]."""
        n = len(order)
        chars = self._chars()
        res = [("", 0.0)] * n
        words = [None] * n
        for (lo, hi, imgW, mx), to in zip(batches, tok_off):
            b = hi - lo
            T = self.engine.tokens(imgW)
            ids = ids_h[to: to + b * T].reshape(b, T)
            pr = probs_h[to: to + b * T].reshape(b, T)
            sel = np.ones((b, T), bool)
            sel[:, 1:] = ids[:, 1:] != ids[:, :-1]
            sel &= ids != 0
            cnt = sel.sum(1)
            flat = pr[sel].astype(np.float64)
            starts = np.concatenate([[0], np.cumsum(cnt)[:-1]])
            conf = np.zeros(b, np.float64)
            nz = cnt > 0
            if flat.size:
                # np.add.reduceat sums each [start, next start) run with numpy's pairwise reduction, i.e. exactly what
                # np.mean does on the run alone; the sentinel keeps the start of an empty trailing run in range, and the
                # last run is summed on its own so that the sentinel never joins a reduction
                sums = np.add.reduceat(np.concatenate([flat, [0.0]]), starts)
                sums[b - 1] = flat[starts[b - 1]:].sum() if cnt[b - 1] else 0.0
                conf[nz] = sums[nz] / cnt[nz]
            conf = conf.round(5)
            kept = ids[sel]
            for j in range(b):
                i = int(order[lo + j])
                seg = kept[starts[j]: starts[j] + cnt[j]]
                text = "".join(chars[seg])
                res[i] = (text, float(conf[j]))
                if return_word_box:
                    wi = get_word_info(text, sel[j]) if text else WordInfo()
                    wi.line_txt_len = T * ratios[i] / mx
                    wi.confs = pr[j][sel[j]].tolist() or [0]
                    words[i] = wi
        return res, words

    def _chars(self):
        if getattr(self, "_char_arr", None) is None:
            self._char_arr = np.array(self.character, dtype=object)
        return self._char_arr


# ----------------------------------------------------------------------------- the model class
class B200OcrModel:
    """Drop-in for `RapidOcrModel` (rapid_doc/model/ocr/rapid_ocr.py:43-162) on one B200."""

    def __init__(self, det_db_box_thresh=0.5, lang=None, ocr_config=None, use_dilation=True, det_db_unclip_ratio=1.8,
                 enable_merge_det_boxes=True, is_seal=False, device=0, precision=None, blobs=None):
        """blobs=(det_blob, rec_blob): packed weights handed in by the caller (multi-GPU: rank 0 packs, NCCL broadcast)."""
        if is_seal:
            raise NotImplementedError("seal OCR (PP-OCRv4 seal det) is outside the B200 hot path; use RapidOcrModel(is_seal=True)")
        cfg = dict(ocr_config or {})
        self.drop_score = 0.5
        self.enable_merge_det_boxes = enable_merge_det_boxes
        self.is_seal = False
        prec = _lib.PREC_FP16 if precision is None else precision
        det = DetEngine(device=device, precision=prec, weights_path=cfg.get("Det.model_path"), blob=blobs[0] if blobs else None)
        rec = RecEngine(device=device, precision=prec, weights_path=cfg.get("Rec.model_path"), blob=blobs[1] if blobs else None)
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
        self.det_window = int(cfg.get("Det.window", 64))      # pages per GPU pass (memory bound only; results do not depend on it)

    # ---- rapid_ocr.py:474-540
    def det_batch_predict(self, img_list, max_batch_size=8):
        """[(boxes, elapse)] per image.  `max_batch_size` bounds the reference's session batch; results here do not depend
        on the batch split (the kernels are batch-invariant), so same-size images are processed as one window, cut only
        for memory (`self.det_window` pages per GPU pass)."""
        if img_list is None or len(img_list) == 0:
            return []
        groups = {}
        for i, im in enumerate(img_list):
            groups.setdefault(tuple(im.shape[:2]), []).append(i)
        out = [None] * len(img_list)
        for idxs in groups.values():
            for c0 in range(0, len(idxs), self.det_window):
                part = idxs[c0: c0 + self.det_window]
                t0 = time.time()
                whole = len(part) == len(img_list)
                res = self.text_detector.detect_batch(img_list if whole else [img_list[i] for i in part])
                el = (time.time() - t0) / len(part)
                for i, (boxes, _) in zip(part, res):
                    if boxes is None:
                        out[i] = (None, 0)
                    else:
                        out[i] = (np.array(sorted_boxes(boxes)) if len(boxes) else boxes, el)
        return out

    def _post_boxes(self, dt_boxes, mfd_res):
        """sorted_boxes -> merge_det_boxes -> update_det_boxes (rapid_ocr.py:262-269, analyze_utils.py:193-203)."""
        dt_boxes = sorted_boxes(dt_boxes)
        if self.enable_merge_det_boxes:
            dt_boxes = merge_det_boxes(dt_boxes)
        if mfd_res:
            dt_boxes = update_det_boxes(dt_boxes, mfd_res)
        return dt_boxes

    # ---- rapid_ocr.py:351-401
    def __call__(self, img, mfd_res=None):
        if img is None:
            return None, None
        res = self.ocr_pages([img], mfd_res_list=[mfd_res])[0]
        if res is None:
            return None, None
        return [np.asarray(b, np.float32) for b, _ in res], [r for _, r in res]

    def _windows(self, pages):
        """Same-size groups cut into windows of det_window pages: [(indices, h, w, sub-batch)]."""
        n = len(pages)
        groups = {}
        for i, im in enumerate(pages):
            groups.setdefault((int(im.shape[0]), int(im.shape[1])), []).append(i)
        out = []
        for (h, w), idxs in groups.items():
            for c0 in range(0, len(idxs), self.det_window):
                part = idxs[c0: c0 + self.det_window]
                if len(part) == n:
                    sub = pages
                elif hasattr(pages, "is_cuda"):
                    sub = pages[part[0]: part[-1] + 1] if part == list(range(part[0], part[-1] + 1)) else pages[part]
                else:
                    sub = [pages[i] for i in part]
                out.append((part, h, w, sub))
        return out

    def _window_crops(self, part, h, w, sub, mfd_res_list):
        """Stage A of one window: det -> sorted / merged boxes -> crop geometry -> every text-line crop warped on the GPU from
        the resident pages.  Returns (boxes per page, DeviceCrops or None, event recorded after the warp)."""
        import torch
        dev_index = self.text_detector.engine.device
        with dbpost.timed("pages.det_window"):
            res, pages_dev = self.text_detector.detect_window(sub, keep_pages=True)
        t_geo = dbpost.timed("pages.sort/merge+crop_geometry")
        t_geo.__enter__()
        boxes_per_page, flat, page_idx = [], [], []
        if not mfd_res_list or not any(mfd_res_list[i] for i in part):
            # sorted_boxes (detector) -> sorted_boxes (caller, rapid_ocr.py:372) -> merge_det_boxes for the whole window in one
            # native call (lines.sort_merge_window == the Python functions below, tests/test_lines.py)
            from .lines import sort_merge_window
            merged = sort_merge_window([b if b is not None else [] for b, _ in res], self.enable_merge_det_boxes)
            for k, bl in enumerate(merged):
                bl = list(bl)
                boxes_per_page.append(bl)
                flat.extend(bl)
                page_idx.extend([k] * len(bl))
        else:
            for k, (i, (boxes, _)) in enumerate(zip(part, res)):
                if boxes is None or len(boxes) == 0:
                    boxes_per_page.append([])
                    continue
                bl = self._post_boxes(np.array(sorted_boxes(boxes)), mfd_res_list[i])
                boxes_per_page.append(bl)
                flat.extend(bl)
                page_idx.extend([k] * len(bl))
        keep, sizes, minv, rot = crop_geometry_batch(flat)
        if len(keep) != len(flat):      # degenerate quads (the reference's warp would fail on them too) are dropped
            alive = set(int(v) for v in keep)
            q, pruned = 0, []
            for bl in boxes_per_page:
                pruned.append([b for j, b in enumerate(bl) if (q + j) in alive])
                q += len(bl)
            boxes_per_page = pruned
            page_idx = [page_idx[int(v)] for v in keep]
        t_geo.__exit__()
        if len(keep) == 0:
            return boxes_per_page, None, None
        pidx = np.asarray(page_idx, np.int32)          # named: the array must outlive the ctypes call
        minv = np.ascontiguousarray(minv)
        sizes = np.ascontiguousarray(sizes)
        nbytes = sizes[:, 0].astype(np.int64) * sizes[:, 1] * 3
        offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
        dev = torch.device("cuda", dev_index)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev)
            buf = torch.empty(int(nbytes.sum()), dtype=torch.uint8, device=dev)
            _lib.check(_lib.load().rdb_warp_crops_batch(dev_index, _lib.ptr(pages_dev), len(part), h, w, len(keep), _lib.ptr(pidx), _lib.ptr(minv),
                                                        _lib.ptr(sizes), _lib.ptr(rot), _lib.ptr(buf), _lib.ptr(offs), int(buf.numel()),
                                                        st.cuda_stream or None))
            ev = torch.cuda.Event()
            ev.record(st)
        self.text_recognizer.stats["launches"] += 1                        # warp_cubic
        self.text_recognizer.stats["h2d_bytes"] += len(keep) * 104
        shapes = np.where(rot[:, None] != 0, sizes, sizes[:, ::-1]).astype(np.int64)      # stored (h, w): rot90 swaps them
        return boxes_per_page, DeviceCrops(buf, offs, shapes, dev_index), ev

    def _window_texts(self, part, boxes_per_page, dc, ev, drop, out):
        """Stage B of one window: one recognition pass over all its crops, results written into out[page index]."""
        import torch
        if dc is None:
            return
        cur = torch.cuda.current_stream(dc.buf.device)
        if ev is not None:
            cur.wait_event(ev)                 # the crops may have been produced on another stream (ocr_pages_stream)
            dc.buf.record_stream(cur)
        with dbpost.timed("pages.rec"):
            rec = self.text_recognizer(dc)
        q = 0
        txts, scores = rec.txts, rec.scores
        for k, i in enumerate(part):
            bl = boxes_per_page[k]
            if len(bl) == 0:
                continue
            quads = np.asarray(bl).tolist()                 # one conversion per page instead of one per box
            page_res = [[quads[j], (txts[q + j], scores[q + j])] for j in range(len(bl)) if scores[q + j] >= drop]
            q += len(bl)
            out[i] = page_res or None

    def ocr_pages(self, pages, mfd_res_list=None, drop_score=None):
        """det + rec for a WINDOW of pages — the fused equivalent of the reference's window flow
        (`_run_ocr_det_batch` analyze_utils.py:105-212 then `_run_ocr_rec_postprocess` :216-292, or `__call__`
        rapid_ocr.py:351-401 per page): det on all pages -> sorted/merged boxes -> every text-line crop of the window
        warped on the GPU from the resident pages (get_rotate_crop_image, bit-exact) -> one recognition pass over all crops.
        Returns per page: None (no boxes) or [[box (4x2 list), (text, score)], ...] with score >= drop_score.
        More than `det_window` pages (or mixed sizes) are processed window by window; each window is its own recognition
        window (its crops are batched together, as the reference batches the crops of one `ocr(det=False)` call)."""
        drop = self.drop_score if drop_score is None else drop_score
        out = [None] * len(pages)
        for part, h, w, sub in self._windows(pages):
            boxes_per_page, dc, ev = self._window_crops(part, h, w, sub, mfd_res_list)
            self._window_texts(part, boxes_per_page, dc, ev, drop, out)
        return out

    def ocr_pages_stream(self, batches, drop_score=None):
        """Generator over an iterable of page batches: yields `ocr_pages(batch)` for each, with the detection stage of the next
        window running on a worker thread (own CUDA stream) while this thread recognises the current one — the host half of
        detection (contours, boxes, crop geometry) hides behind the recogniser's GPU time and vice versa.  Results are
        identical to calling `ocr_pages` on each batch (same windows, same recognition batches)."""
        import torch
        from concurrent.futures import ThreadPoolExecutor
        drop = self.drop_score if drop_score is None else drop_score
        dev = torch.device("cuda", self.text_detector.engine.device)
        if getattr(self, "_stage_a", None) is None:
            self._stage_a = ThreadPoolExecutor(max_workers=1)
            self._stream_a = torch.cuda.Stream(dev, priority=-1)      # detection stage ahead of the running recogniser

        def stage_a(pages):
            with torch.cuda.device(dev), torch.cuda.stream(self._stream_a):
                return [(part, self._window_crops(part, h, w, sub, None)) for part, h, w, sub in self._windows(pages)]
        it = iter(batches)
        try:
            cur_pages = next(it)
        except StopIteration:
            return
        # two Python threads share the GIL: with the default 5 ms switch interval every return from a C call in one stage can
        # wait that long for the other stage's pure-Python loop; a short interval keeps both stages moving
        import sys
        old_switch = sys.getswitchinterval()
        sys.setswitchinterval(min(old_switch, 2e-4))
        fut = self._stage_a.submit(stage_a, cur_pages)
        try:
            yield from self._stream_loop(fut, it, cur_pages, stage_a, drop)
        finally:
            sys.setswitchinterval(old_switch)

    def _stream_loop(self, fut, it, cur_pages, stage_a, drop):
        while fut is not None:
            staged, n_cur = fut.result(), len(cur_pages)
            try:
                cur_pages = next(it)
                fut = self._stage_a.submit(stage_a, cur_pages)
            except StopIteration:
                fut = None
            out = [None] * n_cur
            for part, (boxes_per_page, dc, ev) in staged:
                self._window_texts(part, boxes_per_page, dc, ev, drop, out)
            yield out

    # ---- rapid_ocr.py:225-299
    def ocr(self, img, det=True, rec=True, mfd_res=None, tqdm_enable=False, tqdm_desc="OCR-rec Predict", return_word_box=False,
            ori_img=None, dt_boxes=None):
        assert isinstance(img, (np.ndarray, list, bytes))
        if isinstance(img, list) and det:
            raise ValueError("When input a list of images, det must be false")
        img = check_img(img)
        if det and rec:
            boxes, res = self.__call__(preprocess_image(img), mfd_res=mfd_res)
            if not boxes and not res:
                return [None]
            return [[[np.asarray(b).tolist(), r] for b, r in zip(boxes, res)]]
        if det and not rec:
            d = self.text_detector(preprocess_image(img))
            if d.boxes is None:
                return [None]
            boxes = self._post_boxes(np.array(d.boxes), mfd_res)
            return [[np.asarray(b).tolist() for b in boxes]]
        crops = img if isinstance(img, list) else [preprocess_image(img)]
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
