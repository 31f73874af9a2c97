"""Table-orientation classifier and seal-text detector forward on the B200, behind the reference's
class interfaces:

  B200Orientation          `RapidOrientation`        rapid_doc/model/orientation/rapid_orientation/rapid_orientation.py:30-56
                           (PreProcess list of config.yaml: ResizeImage resize_short 256 -> CropImage 224 -> NormalizeImage ->
                           ToCHWImage, utils.py:97-172; labels from the model file's `character` metadata)
  B200OrientationModel     `RapidOrientationModel`   rapid_doc/model/orientation/rapid_orientation_model.py:7-53
  B200SealDetector         the `Det.*` half of `RapidOcrModel(is_seal=True)`  rapid_doc/model/ocr/rapid_ocr.py:122-143
                           (limit_side_len 736 / limit_type 'min', thresh 0.2; network pp-ocrv4_mobile_seal_det.onnx) up to the
                           text polygons (`detect`: probability map, bitmap, polygon-mode DB post-process); `sort_poly_boxes` =
                           SortPolyBoxes (model/ocr/seal_crop.py:26-39)

The networks run through `onnx_run.OnnxCnn` (CUDA, fp32).  No CPU path.
"""
import os
import time

import cv2
import numpy as np

from .onnx_run import OnnxCnn
from .table import needs_orientation_cls
from .weights import WEIGHTS_DIR

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406]).reshape(1, 1, 3).astype("float32")
IMAGENET_STD = np.array([0.229, 0.224, 0.225]).reshape(1, 1, 3).astype("float32")


def orientation_preprocess(img, resize_short=256, size=224):
    h, w = img.shape[:2]
    pct = float(resize_short) / min(w, h)
    img = cv2.resize(img, (int(round(w * pct)), int(round(h * pct))))
    h, w = img.shape[:2]
    if h < size or w < size:
        raise ValueError(f"The size({size}, {size}) of CropImage must be greater than size({h}, {w}) of image.")
    ws, hs = (w - size) // 2, (h - size) // 2
    img = img[hs:hs + size, ws:ws + size, :]
    x = (np.array(img).astype(np.float32) * np.float32(1.0 / 255.0) - IMAGENET_MEAN) / IMAGENET_STD
    return x.astype(np.float32).transpose((2, 0, 1))


class B200Orientation:
    def __init__(self, model_path=None, device=0):
        if model_path is None:
            model_path = os.path.join(WEIGHTS_DIR, "rapid_orientation.onnx")
        self.session = OnnxCnn(model_path, device)
        self.labels = self.session.meta["character"].splitlines()

    def scores(self, images):
        """A batch of images -> [n, 4] softmax scores (one forward for all of them; the reference runs them one at a time)."""
        x = np.stack([orientation_preprocess(im) for im in images])
        return self.session(x)

    def __call__(self, images):
        s = time.time()
        out = self.scores([images]).squeeze()
        return self.labels[int(np.argmax(out))], time.time() - s


class B200OrientationModel:
    def __init__(self, device=0, model_path=None):
        self.orientation_engine = B200Orientation(model_path, device)

    def predict(self, input_img, det_res=None):
        bgr = cv2.cvtColor(input_img, cv2.COLOR_RGB2BGR)
        if not needs_orientation_cls(bgr.shape, det_res):
            return "0"
        return self.orientation_engine(input_img)[0]


def polygons_from_bitmap(pred, bitmap, dest_width, dest_height, box_thresh=0.6, unclip_ratio=0.5, max_candidates=1000, min_size=3):
    """DB post-process in polygon mode (`Det.box_type = 'poly'`, the seal configuration of rapid_ocr.py:122-131), restated from
    the published PaddleOCR `DBPostProcess.polygons_from_bitmap` — the implementation RapidDoc calls lives in rapidocr (absent
    here): **parity unpinned**.  contour -> approxPolyDP(0.002 * perimeter) -> mean probability inside the polygon ->
    offset by area * unclip_ratio / perimeter (the library's Clipper restatement; pyclipper's final union of self-intersecting
    offsets is not reproduced) -> min-area-rect size filter -> scale to the source image.  -> (list of [k,2] int arrays, scores)"""
    import ctypes as C
    from . import _lib
    from .dbpost import cv2_score_fn
    height, width = bitmap.shape
    contours, _ = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    score_one = cv2_score_fn(None).score_one
    lib = _lib.load()
    boxes, scores = [], []
    for contour in contours[:max_candidates]:
        approx = cv2.approxPolyDP(contour, 0.002 * cv2.arcLength(contour, True), True)
        points = approx.reshape((-1, 2))
        if points.shape[0] < 4:
            continue
        score = score_one(pred, points.astype(np.float32))
        if box_thresh > score:
            continue
        area, length = abs(cv2.contourArea(points.astype(np.float32))), cv2.arcLength(points.astype(np.float32), True)
        if length <= 0:
            continue
        px, py = points[:, 0].astype(np.float64), points[:, 1].astype(np.float64)
        if float(np.sum(px * np.roll(py, -1) - np.roll(px, -1) * py)) < 0:      # Clipper orients closed paths before offsetting
            points = points[::-1]                                                # (same orientation as the ordered mini boxes)
        xy = (C.c_double * (2 * len(points)))(*[float(v) for v in points.reshape(-1)])
        cap = 4096
        out = (C.c_int64 * (2 * cap))()
        n = lib.rdb_clipper_offset(xy, len(points), float(area * unclip_ratio / length), out, cap)
        if n <= 0:
            continue
        box = np.frombuffer(out, dtype=np.int64)[: 2 * n].reshape(-1, 2).astype(np.int32)
        if not cv2.isContourConvex(points.reshape(-1, 1, 2).astype(np.int32)):
            # the raw offset of a concave polygon carries a swallow-tail loop at every reflex vertex; pyclipper removes them with
            # a polygon union.  Here the loops are dropped by re-tracing the outline of the filled offset path united with the
            # source polygon (cv2 raster, so the vertices are pixel-chain corners rather than Clipper's arc points)
            x0, y0 = box.min(axis=0) - 2
            x1, y1 = box.max(axis=0) + 3
            m = np.zeros((int(y1 - y0), int(x1 - x0)), np.uint8)
            cv2.fillPoly(m, [(points.astype(np.int32) - [x0, y0]).reshape(-1, 1, 2)], 1)
            cv2.polylines(m, [(box - [x0, y0]).reshape(-1, 1, 2)], True, 1, 1)
            ff = m.copy()
            mask = np.zeros((m.shape[0] + 2, m.shape[1] + 2), np.uint8)
            cv2.floodFill(ff, mask, (0, 0), 2)                                   # outside of the outer boundary
            outer, _ = cv2.findContours((ff != 2).astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
            if len(outer) != 1:
                continue
            eps = 0.001 * cv2.arcLength(outer[0], True)
            box = (cv2.approxPolyDP(outer[0], eps, True).reshape(-1, 2) + [x0, y0]).astype(np.int32)
        rect = cv2.minAreaRect(box.reshape((-1, 1, 2)))
        if min(rect[1]) < min_size + 2:
            continue
        box = box.astype(np.float64)
        box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_width), 0, dest_width)
        box[:, 1] = np.clip(np.round(box[:, 1] / height * dest_height), 0, dest_height)
        boxes.append(box.astype(np.int32))
        scores.append(float(score))
    return boxes, scores


def sort_poly_boxes(dt_polys):
    """Seal polygons top to bottom by their smallest y (stable argsort)."""
    if len(dt_polys) == 0:
        return dt_polys
    rank = np.argsort(np.array([min(np.asarray(p)[:, 1]) for p in dt_polys]))
    return [dt_polys[i] for i in rank]


class B200SealDetector:
    def __init__(self, device=0, model_path=None, limit_side_len=736, limit_type="min", thresh=0.2):
        if model_path is None:
            model_path = os.path.join(WEIGHTS_DIR, "pp-ocrv4_mobile_seal_det.onnx")
        self.session = OnnxCnn(model_path, device)
        self.limit_side_len, self.limit_type, self.thresh = limit_side_len, limit_type, thresh

    def preprocess(self, img):
        """DetPreProcess of rapidocr as RapidOcrModel configures it for seals: scale the short side up to 736, round to /32,
        (x/255 - mean)/std, CHW."""
        h, w = img.shape[:2]
        if self.limit_type == "max":
            ratio = float(self.limit_side_len) / max(h, w) if max(h, w) > self.limit_side_len else 1.0
        else:
            ratio = float(self.limit_side_len) / min(h, w) if min(h, w) < self.limit_side_len else 1.0
        rh, rw = int(round(int(h * ratio) / 32) * 32), int(round(int(w * ratio) / 32) * 32)
        if rh <= 0 or rw <= 0:
            return None
        img = cv2.resize(img, (rw, rh))
        x = (img.astype("float32") / 255.0 - IMAGENET_MEAN) / IMAGENET_STD
        return np.ascontiguousarray(x.transpose((2, 0, 1))[None], np.float32)

    def prob_map(self, img):
        x = self.preprocess(img)
        return None if x is None else self.session(x)[0, 0]

    def __call__(self, img):
        """-> (probability map [rh, rw] float32, bitmap uint8) at the network resolution."""
        p = self.prob_map(img)
        return p, (p > self.thresh).astype(np.uint8)

    def detect(self, img, box_thresh=0.6, unclip_ratio=0.5):
        """Seal text polygons in image coordinates, top to bottom (the `dt_boxes` RapidOcrModel hands to SortPolyBoxes /
        CropByPolys, rapid_ocr.py:266-268), with their scores."""
        prob, bitmap = self(img)
        polys, scores = polygons_from_bitmap(prob, bitmap, img.shape[1], img.shape[0], box_thresh, unclip_ratio)
        order = np.argsort(np.array([p[:, 1].min() for p in polys])) if polys else []
        return [polys[i] for i in order], [scores[i] for i in order]
