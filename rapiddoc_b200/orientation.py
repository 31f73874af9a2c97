"""Table-orientation classifier and seal-text detector forward on the B200, behind the reference's
class interfaces:

  B200Orientation          `RapidOrientation`        rapid_doc/model/orientation/rapid_orientation/rapid_orientation.py:30-56
                           (PreProcess list of config.yaml: ResizeImage resize_short 256 -> CropImage 224 -> NormalizeImage ->
                           ToCHWImage, utils.py:97-172; labels from the model file's `character` metadata)
  B200OrientationModel     `RapidOrientationModel`   rapid_doc/model/orientation/rapid_orientation_model.py:7-53
  B200SealDetector         the `Det.*` half of `RapidOcrModel(is_seal=True)`  rapid_doc/model/ocr/rapid_ocr.py:122-143
                           (limit_side_len 736 / limit_type 'min', thresh 0.2; network pp-ocrv4_mobile_seal_det.onnx) up to the
                           probability map and bitmap; `sort_poly_boxes` = SortPolyBoxes (model/ocr/seal_crop.py:26-39)

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
