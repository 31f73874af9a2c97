"""Layout path: the numeric work around the PP-DocLayout network, for a WINDOW of pages.

Reference (all under rapid_doc/):
  L1  RapidLayoutModel.batch_predict          model/layout/rapid_layout.py:55-108 (label -> CategoryId maps :131-227)
  L2  PPPreProcess                             model/layout/rapid_layout_self/model_handler/pp_doclayout/pre_process.py:11-42
  L3  the detector itself                      ONNX files that are downloaded at first run — NOT available offline and there is
                                               no network source in the reference: `B200LayoutModel` takes any object with the
                                               reference's InferSession protocol (inference_engine/base.py:15-57) as `session`;
                                               without one it raises "skipped: weights unavailable"
  L4  PPPostProcess                            .../pp_doclayout/post_process.py:20-243, nms :948-979, check_containment :996-1022,
                                               unclip_boxes :611-662, restructured_boxes :567-609
  L5  filter_overlap_boxes                     backend/utils/utils.py:109-173

The two O(n^2) Python loops of L4 (class-aware NMS, containment) run on the GPU for all pages of the window in one launch
each (rdb_layout_nms / rdb_layout_containment, float32 arithmetic in the reference's operation order: keep-sets and flags are
bit-identical); thresholds, sorting by the order column and the dict building are array code on the host.  Masks (-> polygons,
PP-DocLayoutV3 "auto" shape mode, shapely) are not handled: `layout_shape_mode` is "rect", as the reference itself falls back
to when the session returns no masks (pp_doclayout/main.py:66-69).
"""
import cv2
import numpy as np

from . import _lib

# CategoryId (rapid_doc/utils/enum_class.py:90-106)
TITLE, TEXT, ABANDON, IMAGE_BODY, IMAGE_CAPTION, TABLE_BODY, TABLE_CAPTION = 0, 1, 2, 3, 4, 5, 6
EQ_NUMBER, INLINE_EQ, INTERLINE_EQ = 9, 13, 14

_TEXT_LIKE = ("text number abstract content figure_title reference footnote header algorithm footer aside_text reference_content "
              "vertical_text vision_footnote").split()
_IMAGE_LIKE = "image seal chart header_image footer_image".split()


def class_map(labels, markdown_ignore_labels=("number", "footnote", "header", "header_image", "footer", "footer_image", "aside_text")):
    """label -> CategoryId for the PP-DocLayout families (rapid_layout.py:131-227): titles, text-like, image-like, tables,
    formulas; labels in `markdown_ignore_labels` become Abandon."""
    out = {}
    for lb in labels:
        if lb in ("paragraph_title", "doc_title"):
            c = TITLE
        elif lb in _TEXT_LIKE:
            c = TEXT
        elif lb in _IMAGE_LIKE:
            c = IMAGE_BODY
        elif lb == "table":
            c = TABLE_BODY
        elif lb == "table_title":
            c = TABLE_CAPTION
        elif lb == "chart_title":
            c = IMAGE_CAPTION
        elif lb in ("formula", "display_formula"):
            c = INTERLINE_EQ
        elif lb == "inline_formula":
            c = INLINE_EQ
        elif lb == "formula_number":
            c = EQ_NUMBER
        else:
            c = TEXT
        out[lb] = ABANDON if lb in markdown_ignore_labels else c
    return out


class LayoutPreProcess:
    """L2: cv2.resize(INTER_CUBIC) to the model size, (x/255 - mean)/std, CHW, batch dim — float32 out."""

    def __init__(self, img_size, imagenet_norm=False):
        self.size = img_size
        self.mean = np.array([0.485, 0.456, 0.406]) if imagenet_norm else np.array([0, 0, 0])
        self.std = np.array([0.229, 0.224, 0.225]) if imagenet_norm else np.array([1.0, 1.0, 1.0])
        self.scale = 1 / 255.0

    def __call__(self, img):
        if img is None:
            raise ValueError("img is None.")
        rh, rw = self.size
        x = cv2.resize(img, (int(rw), int(rh)), interpolation=2)
        x = (x.astype("float32") * self.scale - self.mean) / self.std
        return np.expand_dims(x.transpose((2, 0, 1)), axis=0).astype(np.float32)


def _flat(pages):
    for b in pages:
        if len(b) and np.asarray(b).dtype != np.float32:
            raise _lib.B200Error("layout boxes must be float32 (the detector session's output dtype): the GPU kernels reproduce the "
                                 "reference's float32 arithmetic bit for bit, other dtypes would follow different rounding")
    offs = np.zeros(len(pages) + 1, np.int32)
    offs[1:] = np.cumsum([len(b) for b in pages])
    if offs[-1] == 0:
        return np.zeros((0, 6), np.float32), offs
    return np.ascontiguousarray(np.concatenate([np.asarray(b, np.float32).reshape(len(b), -1)[:, :6] for b in pages if len(b)]), np.float32), offs


def nms_window(pages, iou_same=0.6, iou_diff=0.98, device=0):
    """`nms(boxes[:, :6], iou_same, iou_diff)` for every page of the window: list of kept index lists (selection order)."""
    flat, offs = _flat(pages)
    if offs[-1] == 0:
        return [[] for _ in pages]
    order = np.concatenate([np.argsort(np.asarray(b)[:, 1])[::-1] if len(b) else np.zeros(0, np.int64) for b in pages]).astype(np.int32)
    keep = np.zeros(int(offs[-1]), np.int32)
    keep_n = np.zeros(len(pages), np.int32)
    _lib.check(_lib.load().rdb_layout_nms(int(device), _lib.ptr(flat), 6, _lib.ptr(order), _lib.ptr(offs), len(pages), float(iou_same), float(iou_diff),
                                          _lib.ptr(keep), _lib.ptr(keep_n), None))
    return [keep[offs[p]: offs[p] + keep_n[p]].tolist() for p in range(len(pages))]


def containment_window(pages, formula_index=None, category_index=None, mode=None, device=0):
    """`check_containment` for every page: list of (contains_other, contained_by_other) int arrays."""
    flat, offs = _flat(pages)
    co = np.zeros(int(offs[-1]), np.int32)
    cb = np.zeros(int(offs[-1]), np.int32)
    if offs[-1]:
        m = {None: 0, "large": 1, "small": 2}[mode]
        _lib.check(_lib.load().rdb_layout_containment(int(device), _lib.ptr(flat), 6, _lib.ptr(offs), len(pages), -1 if formula_index is None else int(formula_index),
                                                      -1 if category_index is None else int(category_index), m, _lib.ptr(co), _lib.ptr(cb), None))
    return [(co[offs[p]: offs[p + 1]].astype(int), cb[offs[p]: offs[p + 1]].astype(int)) for p in range(len(pages))]


class LayoutPostProcess:
    """L4: PPPostProcess.__call__ for a window (rect shape mode)."""

    def __init__(self, labels, conf_thres=0.5, layout_nms=True, layout_merge_bboxes_mode=None, layout_unclip_ratio=None, scale_size=None,
                 device=0):
        self.labels, self.conf_thres, self.layout_nms = labels, conf_thres, layout_nms
        self.merge_mode, self.unclip_ratio, self.scale_size, self.device = layout_merge_bboxes_mode, layout_unclip_ratio, scale_size, device

    def _threshold(self, boxes):
        t = self.conf_thres
        if isinstance(t, float):
            return boxes[(boxes[:, 1] > t) & (boxes[:, 0] > -1), :]
        parts = []
        for cat in np.unique(boxes[:, 0]):
            cb = boxes[boxes[:, 0] == cat]
            parts.append(cb[(cb[:, 1] > t.get(int(cat), 0.5)) & (cb[:, 0] > -1)])
        return np.vstack(parts) if parts else np.array([])

    def _drop_page_sized_images(self, boxes, img_size):
        if not (len(boxes) > 1 and boxes.shape[1] in (6, 7, 8)):
            return boxes
        area_thres = 0.82 if img_size[0] > img_size[1] else 0.93
        image_index = self.labels.index("image") if "image" in self.labels else None
        img_area = img_size[0] * img_size[1]
        keep = []
        for box in boxes:
            label_index, _, xmin, ymin, xmax, ymax = box[:6]
            if label_index == image_index:
                xmin, ymin = max(0, xmin), max(0, ymin)
                xmax, ymax = min(img_size[0], xmax), min(img_size[1], ymax)
                if (xmax - xmin) * (ymax - ymin) <= area_thres * img_area:
                    keep.append(box)
            else:
                keep.append(box)
        return np.array(keep if keep else boxes)

    def __call__(self, boxes_per_page, img_sizes):
        """boxes_per_page: the session's per-page box arrays [n, 6|7|8]; img_sizes: (w, h) per page.
        Returns per page the list of dicts of `restructured_boxes` (or np.array([]) when nothing is left)."""
        pages = [self._threshold(np.asarray(b)) for b in boxes_per_page]
        pages = [p if p.ndim == 2 else np.zeros((0, 6), np.float32) for p in pages]
        if self.layout_nms:
            kept = nms_window(pages, 0.6, 0.98, self.device)
            pages = [np.array(p[k]) if len(p) else p for p, k in zip(pages, kept)]
        pages = [self._drop_page_sized_images(p, s) for p, s in zip(pages, img_sizes)]
        mm = self.merge_mode
        if mm:
            formula_index = self.labels.index("formula") if "formula" in self.labels else None
            if isinstance(mm, str):
                assert mm in ("union", "large", "small")
                if mm != "union":
                    rel = containment_window(pages, formula_index, device=self.device)
                    pages = [p[cb == 0] if mm == "large" else p[(co == 0) | (cb == 1)] for p, (co, cb) in zip(pages, rel)]
            else:
                masks = [np.ones(len(p), dtype=bool) for p in pages]
                for cat, lm in mm.items():
                    assert lm in ("union", "large", "small")
                    if lm == "union":
                        continue
                    rel = containment_window(pages, formula_index, cat, lm, self.device)
                    for m, (co, cb) in zip(masks, rel):
                        m &= (cb == 0) if lm == "large" else ((co == 0) | (cb == 1))
                pages = [p[m] for p, m in zip(pages, masks)]
        out = []
        for boxes, img_size in zip(pages, img_sizes):
            if boxes.size == 0:
                out.append(np.array([]))
                continue
            if boxes.shape[1] == 8:
                boxes = boxes[np.lexsort((-boxes[:, 7], boxes[:, 6]))][:, :6]
            if boxes.shape[1] == 7:
                boxes = boxes[np.argsort(boxes[:, 6])][:, :6]
            ur = self.unclip_ratio
            if ur:
                if isinstance(ur, float):
                    ur = (ur, ur)
                boxes = unclip_boxes(boxes, ur)
            if boxes.shape[1] != 6:
                raise ValueError(f"The shape of boxes should be 6 or 10, instead of {boxes.shape[1]}")
            out.append(restructured_boxes(boxes, self.labels, img_size))
        return out


def unclip_boxes(boxes, unclip_ratio):
    """post_process.py:611-662 (tuple ratio and per-class dict)."""
    if isinstance(unclip_ratio, dict):
        rows = []
        for box in boxes:
            cid, score, x1, y1, x2, y2 = box
            if cid in unclip_ratio:
                wr, hr = unclip_ratio[cid]
                w, h = x2 - x1, y2 - y1
                cx, cy = x1 + w / 2, y1 + h / 2
                rows.append([cid, score, cx - w * wr / 2, cy - h * hr / 2, cx + w * wr / 2, cy + h * hr / 2])
            else:
                rows.append(box)
        return np.array(rows)
    w = boxes[:, 4] - boxes[:, 2]
    h = boxes[:, 5] - boxes[:, 3]
    nw, nh = w * unclip_ratio[0], h * unclip_ratio[1]
    cx, cy = boxes[:, 2] + w / 2, boxes[:, 3] + h / 2
    return np.column_stack((boxes[:, 0], boxes[:, 1], cx - nw / 2, cy - nh / 2, cx + nw / 2, cy + nh / 2))


def restructured_boxes(boxes, labels, img_size):
    """post_process.py:567-609: clip to the page, drop empty boxes, number them."""
    w, h = img_size
    out = []
    for idx, box in enumerate(boxes):
        xmin, ymin, xmax, ymax = box[2:]
        xmin, ymin = float(max(0, xmin)), float(max(0, ymin))
        xmax, ymax = float(min(w, xmax)), float(min(h, ymax))
        if xmax <= xmin or ymax <= ymin:
            continue
        out.append({"cls_id": int(box[0]), "label": labels[int(box[0])], "score": float(box[1]), "coordinate": [xmin, ymin, xmax, ymax],
                    "order": idx + 1})
    return out


def _overlap_small(b1, b2):
    """calculate_overlap_ratio(b1, b2, "small") (rapid_doc/model/reading_order/utils.py:10-50): intersection over the smaller area."""
    iw = max(0, min(b1[2], b2[2]) - max(b1[0], b2[0]))
    ih = max(0, min(b1[3], b2[3]) - max(b1[1], b2[1]))
    inter = float(iw) * float(ih)
    ref = min(_area(b1), _area(b2))
    return 0.0 if ref == 0 else inter / ref


def _area(b):
    x1, y1, x2, y2 = map(float, b)
    return abs((x2 - x1) * (y2 - y1))


def filter_overlap_boxes(layout_det_res, use_custom_ocr=False):
    """L5 (backend/utils/utils.py:109-173) for rect results (no polygon_points): drop tiny boxes and the smaller of two boxes
    overlapping by more than 0.7 of the smaller one (inline formulas and image/seal/chart pairs are special-cased)."""
    from copy import deepcopy
    boxes = [b for b in deepcopy(layout_det_res) if b["original_label"] != "reference"]
    dropped = set()
    for i in range(len(boxes)):
        ci = [boxes[i]["poly"][0], boxes[i]["poly"][1], boxes[i]["poly"][4], boxes[i]["poly"][5]]
        if ci[2] - ci[0] < 6 or ci[3] - ci[1] < 6:
            dropped.add(i)
        for j in range(i + 1, len(boxes)):
            if i in dropped or j in dropped:
                continue
            cj = [boxes[j]["poly"][0], boxes[j]["poly"][1], boxes[j]["poly"][4], boxes[j]["poly"][5]]
            ratio = _overlap_small(ci, cj)
            li, lj = boxes[i]["original_label"], boxes[j]["original_label"]
            if li == "inline_formula" or lj == "inline_formula":
                if not use_custom_ocr:
                    continue
                if ratio > 0.5:
                    if li == "inline_formula":
                        dropped.add(i)
                    if lj == "inline_formula":
                        dropped.add(j)
                    continue
            if ratio > 0.7:
                if boxes[i].get("polygon_points"):
                    raise NotImplementedError("polygon overlap (PP-DocLayoutV3 masks, shapely) is outside the B200 layout path")
                if {li, lj} & {"image", "seal", "chart"} and li != lj:
                    continue
                dropped.add(j if _area(ci) >= _area(cj) else i)
    return [b for k, b in enumerate(boxes) if k not in dropped]


MODEL_SIZES = {"pp_doclayout_plus_l": (800, 800), "pp_doclayoutv2": (800, 800), "pp_doclayoutv3": (800, 800), "pp_doclayout_s": (480, 480)}


class B200LayoutModel:
    """L1: `RapidLayoutModel.batch_predict(images, batch_size) -> list[list[dict]]` (rapid_layout.py:55-108) around a
    caller-supplied detector session.  `session(img_inputs [B,3,S,S] f32, scale_factor [B,2] f32) -> [boxes, box_nums]`
    is the reference's InferSession contract (per-model I/O); the network weights are not available offline."""

    def __init__(self, session=None, model_type="pp_doclayoutv3", labels=None, conf_thres=0.3, layout_merge_bboxes_mode=None,
                 markdown_ignore_labels=None, device=0):
        if session is None:
            raise _lib.B200Error("skipped: weights unavailable — the PP-DocLayout detector is an ONNX file RapidDoc downloads at first run; "
                                 "pass an InferSession-protocol object as `session`")
        assert labels, "labels: the class names of the detector"
        self.session, self.model_type, self.labels = session, model_type, list(labels)
        self.img_size = MODEL_SIZES.get(model_type, (640, 640))
        big = model_type in ("pp_doclayout_l", "pp_doclayout_plus_l", "pp_doclayoutv2", "pp_doclayoutv3")
        self.pre = LayoutPreProcess(self.img_size, imagenet_norm=not big)
        ur = [1.0, 1.0] if model_type in ("pp_doclayout_plus_l", "pp_doclayoutv2", "pp_doclayoutv3") else None
        self.post = LayoutPostProcess(self.labels, conf_thres, True, layout_merge_bboxes_mode, ur, self.img_size, device)
        self.cls = class_map(self.labels, markdown_ignore_labels) if markdown_ignore_labels is not None else class_map(self.labels)
        self.ordered = model_type in ("pp_doclayoutv2", "pp_doclayoutv3")

    def predict(self, image):
        return self.batch_predict([image], 1)[0]

    def batch_predict(self, images, batch_size=1):
        res = []
        for b0 in range(0, len(images), max(1, batch_size)):
            chunk = images[b0: b0 + max(1, batch_size)]
            x = np.concatenate([self.pre(im) for im in chunk], axis=0)
            sf = np.array([[self.img_size[0] / im.shape[0], self.img_size[1] / im.shape[1]] for im in chunk], np.float32)
            pred = self.session(x, sf)
            boxes, nums = pred[0], pred[1]
            per, s = [], 0
            for k in range(len(nums)):
                per.append(np.array(boxes[s: s + int(nums[k])]))
                s += int(nums[k])
            datas = self.post(per, [(im.shape[1], im.shape[0]) for im in chunk])
            for d in datas:
                page = []
                for order, item in enumerate(d if len(d) else []):
                    xmin, ymin, xmax, ymax = item["coordinate"]
                    page.append({"category_id": self.cls[item["label"]], "original_label": item["label"],
                                 "original_order": order if self.ordered else -1,
                                 "poly": [xmin, ymin, xmax, ymin, xmax, ymax, xmin, ymax], "polygon_points": None,
                                 "score": round(float(item["score"]), 3)})
                if not self.ordered:
                    for it in page:           # check_inline_formula (rapid_layout.py:110-122)
                        if it["category_id"] == INTERLINE_EQ:
                            bx = (it["poly"][0], it["poly"][1], it["poly"][4], it["poly"][5])
                            for ot in page:
                                if ot["category_id"] == TEXT and _iou(bx, (ot["poly"][0], ot["poly"][1], ot["poly"][4], ot["poly"][5])) >= 0.9:
                                    it["category_id"] = INLINE_EQ
                                    break
                res.append(page)
        return res


def _iou(b1, b2):
    """calculate_iou (rapid_doc/utils/boxbase.py): plain IoU of two xyxy boxes."""
    x1, y1, x2, y2 = max(b1[0], b2[0]), max(b1[1], b2[1]), min(b1[2], b2[2]), min(b1[3], b2[3])
    if x2 < x1 or y2 < y1:
        return 0.0
    inter = (x2 - x1) * (y2 - y1)
    a1, a2 = (b1[2] - b1[0]) * (b1[3] - b1[1]), (b2[2] - b2[0]) * (b2[3] - b2[1])
    if a1 == 0 or a2 == 0:
        return 0.0
    return inter / float(a1 + a2 - inter)
