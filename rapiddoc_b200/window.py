"""The two OCR callers of RapidDoc's batch scheduler, restated so that a window of pages can be run
against `B200OcrModel` / `B200OcrPool` without an importable `rapid_doc` package — and as drop-in replacements for them
(`plugin.install_window()` rebinds the reference's functions to these):

  run_ocr_det_batch        `_run_ocr_det_batch`         rapid_doc/backend/pipeline/analyze_utils.py:105-212
  run_ocr_rec_postprocess  `_run_ocr_rec_postprocess`   rapid_doc/backend/pipeline/analyze_utils.py:216-292
  helpers                  crop_img utils/model_utils.py:90-124; get_adjusted_mfdetrec_res utils/ocr_utils.py:320-342;
                           get_ocr_result_list utils/ocr_utils.py:361-432 (is_mostly_tilted :344-359);
                           _apply_mask_boxes_to_image analyze_utils.py:81-103 (normalize_to_int_bbox utils/bbox_utils.py:6-49)

Same data contract: `ocr_res_all_page` = per page {'ocr_res_list', 'ocr_enable', 'np_img' (RGB), 'single_page_mfdetrec_res',
'checkbox_res', 'lang', 'layout_res'}; detection appends category-15 span dicts (with 'np_img' crops) to 'layout_res',
recognition fills 'text' / 'score' and relabels low-confidence spans.  What changes is only where the time goes: all blocks of
one 64-px size bucket go to the detector as ONE window (the reference passes `Det.rec_batch_num`, default 1), and the
recogniser receives every crop of the window in one call, which `B200TextRecognizer` turns into one upload + one packing
launch + back-to-back batches.  tests/test_callers.py runs these against the reference's real functions (imported by path)
with a deterministic stand-in model.
"""
import copy
import math
from collections import defaultdict

import cv2
import numpy as np

from .lines import merge_det_boxes, sorted_boxes, update_det_boxes
from .ocr import get_rotate_crop_image

OCR_TEXT, LOW_SCORE_TEXT = 15, 16          # CategoryId.OcrText / LowScoreText (utils/enum_class.py:103-104)
MIN_CONFIDENCE, MIN_WIDTH = 0.5, 3         # OcrConfidence (utils/ocr_utils.py:9-11)
RESOLUTION_GROUP_STRIDE = 64


def crop_img(res, img, paste_x=0, paste_y=0, layout_shape_mode="auto"):
    x0, y0, x1, y1 = int(res["poly"][0]), int(res["poly"][1]), int(res["poly"][4]), int(res["poly"][5])
    nw, nh = x1 - x0 + paste_x * 2, y1 - y0 + paste_y * 2
    out = np.ones((nh, nw, 3), dtype=np.uint8) * 255
    crop = img[y0:y1, x0:x1]
    polygon = res.get("polygon_points")
    if layout_shape_mode != "rect" and polygon:
        polygon = np.array(polygon, dtype=np.int32)
        if polygon.ndim == 1:
            polygon = polygon.reshape((-1, 2))
        polygon = polygon.reshape((-1, 1, 2)) - np.array([x0, y0])
        mask = np.zeros(crop.shape[:2], dtype=np.uint8)
        cv2.fillPoly(mask, [polygon], 1)
        crop = crop.copy()
        crop[~mask.astype(bool)] = 255
    out[paste_y:paste_y + (y1 - y0), paste_x:paste_x + (x1 - x0)] = crop
    return out, [paste_x, paste_y, x0, y0, x1, y1, nw, nh]


def adjusted_mfdetrec_res(mfd_res, useful_list):
    px, py, xmin, ymin, _xmax, _ymax, nw, nh = useful_list
    out = []
    for mf in mfd_res:
        a, b, c, d = mf["bbox"]
        x0, y0, x1, y1 = a - xmin + px, b - ymin + py, c - xmin + px, d - ymin + py
        if any([x1 < 0, y1 < 0]) or any([x0 > nw, y0 > nh]):
            continue
        out.append({"bbox": [x0, y0, x1, y1]})
    return out


def _int_bbox(box, image_size):
    arr = np.asarray(box, dtype=np.float64)
    if arr.size == 0:
        return None
    if arr.ndim == 2 and arr.shape[-1] == 2:
        xmin, ymin, xmax, ymax = float(arr[:, 0].min()), float(arr[:, 1].min()), float(arr[:, 0].max()), float(arr[:, 1].max())
    else:
        flat = arr.reshape(-1)
        if flat.size == 4:
            xmin, ymin, xmax, ymax = [float(v) for v in flat]
        elif flat.size >= 8:
            xmin, ymin, xmax, ymax = float(flat[0::2].min()), float(flat[1::2].min()), float(flat[0::2].max()), float(flat[1::2].max())
        else:
            return None
    xmin, ymin, xmax, ymax = math.floor(xmin), math.floor(ymin), math.ceil(xmax), math.ceil(ymax)
    h, w = image_size
    xmin, xmax = max(0, min(int(w), xmin)), max(0, min(int(w), xmax))
    ymin, ymax = max(0, min(int(h), ymin)), max(0, min(int(h), ymax))
    if xmax <= xmin or ymax <= ymin:
        return None
    return [int(xmin), int(ymin), int(xmax), int(ymax)]


def mask_boxes(bgr, boxes):
    """Formula / checkbox regions are painted white before detection."""
    if not boxes:
        return bgr
    out = bgr.copy()
    for mb in boxes:
        bb = mb.get("bbox")
        ib = _int_bbox(bb, out.shape[:2]) if bb is not None else None
        if ib is not None:
            out[ib[1]:ib[3], ib[0]:ib[2]] = 255
    return out


def _is_angle(poly):
    p1, p2, p3, p4 = poly
    height = ((p4[1] - p1[1]) + (p3[1] - p2[1])) / 2
    return not (0.8 * height <= (p3[1] - p1[1]) <= 1.2 * height)


def _mostly_tilted(ocr_res, threshold=1.0):
    angles = [round(abs(math.degrees(math.atan2(p[1][1] - p[0][1], p[1][0] - p[0][0]))) % 180, 2) for p in ocr_res]
    if not angles:
        return False
    avg = round(sum(angles) / len(angles), 2)
    return abs(avg) > threshold and abs(avg - 180) > threshold


def ocr_result_list(ocr_res, useful_list, ocr_enable, bgr_image, lang, original_label, original_order=-1):
    """Detected quads of one block -> span dicts in page coordinates (+ the text-line crop when OCR is enabled)."""
    if not ocr_enable and _mostly_tilted(ocr_res):
        ocr_enable = True
    px, py, xmin, ymin, _xmax, _ymax, _nw, _nh = useful_list
    out = []
    ori = bgr_image.copy()
    for item in ocr_res:
        if len(item) == 2:
            p1, p2, p3, p4 = item[0]
            text, score = item[1]
            if score < MIN_CONFIDENCE:
                continue
        else:
            p1, p2, p3, p4 = item
            text, score = "", 1
            if ocr_enable:
                crop = get_rotate_crop_image(ori, copy.deepcopy(np.array([p1, p2, p3, p4]).astype("float32")))
        poly = [p1, p2, p3, p4]
        if (p3[0] - p1[0]) < MIN_WIDTH:
            continue
        if _is_angle(poly):
            xc, yc = sum(p[0] for p in poly) / 4, sum(p[1] for p in poly) / 4
            nh, nw = ((p4[1] - p1[1]) + (p3[1] - p2[1])) / 2, p3[0] - p1[0]
            p1, p2 = [xc - nw / 2, yc - nh / 2], [xc + nw / 2, yc - nh / 2]
            p3, p4 = [xc + nw / 2, yc + nh / 2], [xc - nw / 2, yc + nh / 2]
        p1, p2, p3, p4 = ([float(p[0] - px + xmin), float(p[1] - py + ymin)] for p in (p1, p2, p3, p4))
        d = {"category_id": OCR_TEXT, "original_label": original_label, "original_order": original_order, "poly": p1 + p2 + p3 + p4}
        if ocr_enable:
            d.update(score=1, text=text, np_img=crop, lang=lang)
        else:
            d.update(score=float(round(score, 2)), text=text)
        out.append(d)
    return out


def run_ocr_det_batch(ocr_res_all_page, get_model, ocr_config):
    """get_model(lang) -> the OCR model object (what AtomModelSingleton.get_atom_model(OCR, lang=...) returns)."""
    use_det_mode = ocr_config.get("use_det_mode", "auto")
    base_batch = ocr_config.get("Det.rec_batch_num", 1)
    infos = []
    for page in ocr_res_all_page:
        for res in page["ocr_res_list"]:
            ocr_enable = page["ocr_enable"]
            if not page["ocr_enable"]:
                if res.get("need_ocr_det"):
                    ocr_enable = True
                elif use_det_mode == "txt" or (use_det_mode != "ocr" and not res.get("need_ocr_det")):
                    continue
            res.pop("need_ocr_det", None)
            new_image, useful = crop_img(res, page["np_img"], 50, 50)
            mfd = adjusted_mfdetrec_res(page["single_page_mfdetrec_res"] + page["checkbox_res"], useful)
            bgr = cv2.cvtColor(new_image, cv2.COLOR_RGB2BGR)
            infos.append((bgr, mask_boxes(bgr, mfd), useful, page, mfd, page["lang"], res, ocr_enable))
    if not infos:
        return
    by_lang = defaultdict(list)
    for info in infos:
        by_lang[info[5]].append(info)
    for lang, group in by_lang.items():
        model = get_model(lang)
        buckets = defaultdict(list)
        for info in group:
            h, w = info[1].shape[:2]
            s = RESOLUTION_GROUP_STRIDE
            buckets[((h + s - 1) // s * s, (w + s - 1) // s * s)].append(info)
        for (th, tw), crops in buckets.items():
            batch = []
            for info in crops:
                h, w = info[1].shape[:2]
                padded = np.ones((th, tw, 3), dtype=np.uint8) * 255
                padded[:h, :w] = info[1]
                batch.append(padded)
            results = model.det_batch_predict(batch, min(len(batch), base_batch))
            for info, (dt_boxes, _) in zip(crops, results):
                bgr, _det, useful, page, mfd, _lang, res, ocr_enable = info
                if dt_boxes is not None and len(dt_boxes) > 0:
                    boxes = sorted_boxes(dt_boxes)
                    boxes = merge_det_boxes(boxes) if boxes else []
                    boxes = update_det_boxes(boxes, mfd) if boxes and mfd else boxes
                    if boxes:
                        quads = [b.tolist() if hasattr(b, "tolist") else b for b in boxes]
                        page["layout_res"].extend(ocr_result_list(quads, useful, ocr_enable, bgr, _lang, res["original_label"], res["original_order"]))


def run_ocr_rec_postprocess(images_layout_res, get_model, ocr_config=None):
    need, crops = {}, {}
    for layout_res in images_layout_res:
        for item in layout_res:
            if item["category_id"] == OCR_TEXT and "np_img" in item and "lang" in item:
                lang = item["lang"]
                need.setdefault(lang, []).append(item)
                crops.setdefault(lang, []).append(item.pop("np_img"))
                item.pop("lang")
    for lang, imgs in crops.items():
        if not imgs:
            continue
        model = get_model(lang)
        try:
            res = model.ocr(imgs, det=False, tqdm_enable=True)[0]
        except Exception:
            res, safe = [], []
            for item, im in zip(need[lang], imgs):
                try:
                    one = model.ocr([im], det=False, tqdm_enable=False)[0]
                except Exception:
                    one = None
                if not one:
                    item["text"], item["score"], item["category_id"] = "", 0.0, LOW_SCORE_TEXT
                    continue
                res.append(one[0])
                safe.append(item)
            need[lang] = safe
        assert len(res) == len(need[lang])
        for item, (text, score) in zip(need[lang], res):
            item["text"] = text
            item["score"] = float(f"{score:.3f}")
            if score < MIN_CONFIDENCE:
                item["category_id"] = LOW_SCORE_TEXT
            else:
                width = item["poly"][4] - item["poly"][0]
                height = item["poly"][5] - item["poly"][1]
                if text in ["（204号", "（20", "（2", "（2号", "（20号", "号", "（204"] and score < 0.8 and width < height:
                    item["category_id"] = LOW_SCORE_TEXT
