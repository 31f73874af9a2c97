"""Table cell / OCR matching and HTML assembly (wireless-table half) and the SLANet orchestration of `RapidTable`
(SURVEY T1), restated so that the table path runs without an importable `rapid_doc`:

  TableMatch           rapid_table_self/table_matcher/main.py:120-348  (filter_ocr_result, match_result with the
                       (1 - IoU, distance) ordering of :75-117, get_pred_html, decode_logic_points)
  format_ocr_results   rapid_table_self/utils/utils.py:15-27
  B200RapidTable       RapidTable.__call__ for the PP-structure models, rapid_table_self/main.py:78-123
  B200RapidTableModel  RapidTableModel.predict / batch_predict for model_type slanet_1m, rapid_doc/model/table/rapid_table.py:110-285
                       (+ normalize_table_ocr_text / _cell_text / _html_cell_text of rapid_doc/model/table/utils.py:22-66)

Host glue (a few hundred boxes per table); the structure model under it is `table.B200TableStructurer` (CUDA).
tests/test_table_match.py runs it against the reference's own class imported by path.
"""
import html as _html
import re
import time

import numpy as np

MIN_IOU = 0.1 ** 8

# ---- OCR-text / cell-text normalisation of the table path (rapid_doc/model/table/utils.py:7-66) ---------------------------
_SINGLE_CHAR = {"香": "否", "哦樂": "哦"}
_REGEX_REPLACEMENTS = ((re.compile(r"^([0-9])號$"), r"\1"),)
_CJK = re.compile(r"[\u3400-\u9fff]")
_CJK_PUNCT = r"，。、“”‘’；：？！、：（）《》【】"


def normalize_table_ocr_text(text):
    """OCR text before table matching: strip, two known single-string repairs, `<digit>號` -> digit, HTML-escape."""
    if text is None:
        return ""
    if not isinstance(text, str):
        text = str(text)
    text = text.strip()
    text = _SINGLE_CHAR.get(text, text)
    for pattern, repl in _REGEX_REPLACEMENTS:
        m = pattern.fullmatch(text)
        if m:
            text = m.expand(repl)
            break
    return _html.escape(text)


def normalize_table_cell_text(text):
    """Spaces the recogniser puts inside Chinese cell text are removed (between CJK characters, around CJK punctuation, between
    CJK and latin / digits)."""
    if not text or not _CJK.search(text):
        return text
    text = re.sub(r"(?<=[\u3400-\u9fff])\s+(?=[\u3400-\u9fff])", "", text)
    text = re.sub(rf"(?<=[\u3400-\u9fffA-Za-z0-9$])\s+(?=[{_CJK_PUNCT}])", "", text)
    text = re.sub(rf"(?<=[{_CJK_PUNCT}])\s+(?=[\u3400-\u9fffA-Za-z0-9$])", "", text)
    text = re.sub(r"(?<=[A-Za-z0-9$])\s+(?=[\u3400-\u9fff])", "", text)
    text = re.sub(r"(?<=[\u3400-\u9fff])\s+(?=[A-Za-z0-9$])", "", text)
    return text


def normalize_table_html_cell_text(html_code):
    """`normalize_table_cell_text` on the text nodes directly inside <td> / <th>.  The reference walks a BeautifulSoup tree
    (bs4 is absent here: **parity unpinned**); the HTML this path produces is flat (`<td ...>text</td>`, optional <b>), so the
    direct children are found with a tag tokenizer.  When nothing changes the input string is returned as it is, as upstream."""
    if not html_code:
        return html_code
    out, depth_cell, changed = [], 0, False
    for tok in re.split(r"(<[^<>]*>)", html_code):
        if tok.startswith("<") and tok.endswith(">"):
            name = re.match(r"</?\s*([a-zA-Z0-9]+)", tok)
            tag = name.group(1).lower() if name else ""
            if tag in ("td", "th"):
                depth_cell = 0 if tok.startswith("</") else 1
            elif depth_cell:
                depth_cell += -1 if tok.startswith("</") else (0 if tok.endswith("/>") else 1)
            out.append(tok)
            continue
        if depth_cell == 1 and tok:
            new = _html.escape(normalize_table_cell_text(_html.unescape(tok)), quote=False)
            if normalize_table_cell_text(_html.unescape(tok)) != _html.unescape(tok):
                changed = True
                tok = new
        out.append(tok)
    return "".join(out) if changed else html_code


def format_ocr_results(ocr_results, img_h, img_w):
    rec_res = list(zip(ocr_results[1], ocr_results[2]))
    boxes = np.array(ocr_results[0])
    lo = np.maximum(boxes[..., :2].min(axis=1), 0)
    hi = np.minimum(boxes[..., :2].max(axis=1), [img_w, img_h])
    return np.hstack([lo, hi]), rec_res


def _cells_xyxy(cell_bboxes):
    """4- or 8-value cell boxes -> [k, 4] float64 (x0, y0, x1, y1)."""
    if cell_bboxes is None:
        return np.empty((0, 4), np.float64)
    rows = []
    for cb in cell_bboxes:
        b = np.asarray(cb, np.float64).reshape(-1)
        if b.size == 8:
            rows.append([b[0::2].min(), b[1::2].min(), b[0::2].max(), b[1::2].max()])
        elif b.size == 4:
            rows.append(b.tolist())
        else:
            raise ValueError(f"Unsupported table cell bbox shape: {b.shape}")
    return np.asarray(rows, np.float64).reshape(-1, 4)


def match_cells(cell_bboxes, dt_boxes, min_iou=MIN_IOU):
    """For every OCR box the cell with the largest IoU (ties: smallest corner distance, then lowest index); boxes whose best IoU
    is below `min_iou` stay unmatched.  -> {cell index: [ocr indices in order]}"""
    dt = np.asarray(dt_boxes, np.float64)
    if dt.size == 0:
        return {}
    dt = dt.reshape(-1, 4)[:, None, :]
    cells = _cells_xyxy(cell_bboxes)
    if cells.size == 0:
        return {}
    c = cells[None, :, :]
    area = (dt[..., 2] - dt[..., 0]) * (dt[..., 3] - dt[..., 1]) + (c[..., 2] - c[..., 0]) * (c[..., 3] - c[..., 1])
    # (the reference names these left/right/top/bottom the other way round; the arithmetic is symmetric)
    y_lo, y_hi = np.maximum(dt[..., 1], c[..., 1]), np.minimum(dt[..., 3], c[..., 3])
    x_lo, x_hi = np.maximum(dt[..., 0], c[..., 0]), np.minimum(dt[..., 2], c[..., 2])
    inter = (y_hi - y_lo) * (x_hi - x_lo)
    hit = (y_lo < y_hi) & (x_lo < x_hi)
    union = area - inter
    iou = np.zeros_like(inter)
    np.divide(inter, union, out=iou, where=hit & (union != 0))
    d_lo = np.abs(c[..., 0] - dt[..., 0]) + np.abs(c[..., 1] - dt[..., 1])
    d_hi = np.abs(c[..., 2] - dt[..., 2]) + np.abs(c[..., 3] - dt[..., 3])
    dist = (d_lo + d_hi) + np.minimum(d_lo, d_hi)
    inv = 1.0 - iou
    matched = {}
    for i in range(inv.shape[0]):
        cand = np.flatnonzero(inv[i] == inv[i].min())
        best = int(cand[np.flatnonzero(dist[i, cand] == dist[i, cand].min())[0]])
        if inv[i, best] >= 1 - min_iou:
            continue
        matched.setdefault(best, []).append(i)
    return matched


class TableMatch:
    def __call__(self, pred_structures, cell_bboxes, dt_boxes, rec_reses):
        out = []
        for struct, cells, dt, rec in zip(pred_structures, cell_bboxes, dt_boxes, rec_reses):
            out.append(None if dt is None or rec is None else self.process_one(struct, cells, dt, rec))
        return out

    def process_one(self, pred_struct, cell_bboxes, dt_boxes, rec_res):
        dt_boxes, rec_res = self.filter_ocr_result(cell_bboxes, dt_boxes, rec_res)
        return self.get_pred_html(pred_struct[0], self.match_result(cell_bboxes, dt_boxes), rec_res)[0]

    @staticmethod
    def filter_ocr_result(cell_bboxes, dt_boxes, rec_res):
        """OCR boxes that end above the first cell are dropped."""
        top = cell_bboxes[:, 1::2].min()
        keep = [i for i, box in enumerate(dt_boxes) if not np.max(box[1::2]) < top]
        return np.array([dt_boxes[i] for i in keep]), [rec_res[i] for i in keep]

    @staticmethod
    def match_result(cell_bboxes, dt_boxes, min_iou=MIN_IOU):
        return match_cells(cell_bboxes, dt_boxes, min_iou)

    @staticmethod
    def get_pred_html(pred_structures, matched_index, ocr_contents):
        html, td = [], 0
        for tag in pred_structures:
            if "</td>" not in tag:
                html.append(tag)
                continue
            if tag == "<td></td>":
                html.append("<td>")
            if td in matched_index:
                idx = matched_index[td]
                many = len(idx) > 1
                bold = many and "<b>" in ocr_contents[idx[0]][0]
                if bold:
                    html.append("<b>")
                parts = []
                for k, j in enumerate(idx):
                    text = ocr_contents[j][0]
                    if many:
                        if len(text) == 0:
                            continue
                        if text[0] == " ":
                            text = text[1:]
                        text = text.replace("<b>", "").replace("</b>", "").strip()
                        if len(text) == 0:
                            continue
                        if k != len(idx) - 1 and text.endswith(" "):
                            text = text.rstrip()
                    parts.append(text)
                html.append(" ".join(parts))
                if bold:
                    html.append("</b>")
            html.append("</td>" if tag == "<td></td>" else tag)
            td += 1
        html = [v for v in html if v not in ("<thead>", "</thead>", "<tbody>", "</tbody>")]
        return "".join(html), html

    def decode_logic_points(self, pred_structures):
        return [np.array(self.decode_one_logic_points(s[0])) for s in pred_structures]

    @staticmethod
    def decode_one_logic_points(tokens):
        """[row_start, row_end, col_start, col_end] of every cell, honouring rowspan / colspan occupancy."""
        points, taken = [], set()
        row = col = 0
        i = 0
        while i < len(tokens):
            tok = tokens[i]
            if tok == "<tr>":
                col = 0
            elif tok == "</tr>":
                row += 1
            elif tok.startswith("<td"):
                cs = rs = 1
                j = i
                if tok != "<td></td>":
                    j += 1
                    while j < len(tokens) and not tokens[j].startswith(">"):
                        if "colspan=" in tokens[j]:
                            cs = int(tokens[j].split("=")[1].strip("\"'"))
                        elif "rowspan=" in tokens[j]:
                            rs = int(tokens[j].split("=")[1].strip("\"'"))
                        j += 1
                i = j
                while (row, col) in taken:
                    col += 1
                points.append([row, row + rs - 1, col, col + cs - 1])
                taken.update((r, c) for r in range(row, row + rs) for c in range(col, col + cs))
                col += cs
            i += 1
        return points


def points_to_bbox(points):
    (x0, y0), (x1, _), (_, y1) = points[0], points[1], points[2]
    return [x0, y0, x1, y1]


def bbox_to_points(bbox):
    x0, y0, x1, y1 = bbox
    return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]]).astype("float32")


class RapidTableOutput:
    def __init__(self):
        self.imgs, self.pred_htmls, self.cell_bboxes, self.logic_points, self.elapse = [], [], [], [], 0.0


class B200RapidTable:
    """`RapidTable(model_type=slanet_1m)` with OCR results supplied by the caller (what RapidTableModel passes,
    rapid_doc/model/table/rapid_table.py:253-262).  All crops of one call go through the structure model as ONE batch — the
    reference's `batch_size` only chunks the Python loop."""

    def __init__(self, device=0, model_path=None, model_type="slanet_1m"):
        from .table import B200TableStructurer
        self.table_structure = B200TableStructurer(model_path, model_type, device)
        self.table_matcher = TableMatch()

    def __call__(self, img_contents, ocr_results=None, batch_size=1, tqdm_enable=False):
        t0 = time.perf_counter()
        imgs = img_contents if isinstance(img_contents, list) else [img_contents]
        for im in imgs:
            if not isinstance(im, np.ndarray):
                raise TypeError(f"Type Error: Expected input of type [ndarray], but received type {type(im).__name__ if im is not None else 'None'}.")
        res = RapidTableOutput()
        structs, cells = self.table_structure(imgs)
        if ocr_results is not None and len(ocr_results) != len(imgs):
            raise ValueError(f"Batch size mismatch: {len(imgs)} images but {len(ocr_results)} OCR results (indices 0:{len(imgs)}).")
        dt_boxes, rec_res = [], []
        for im, ocr in zip(imgs, ocr_results or []):
            d, r = format_ocr_results(ocr, *im.shape[:2])
            dt_boxes.append(d)
            rec_res.append(r)
        res.imgs.extend(imgs)
        res.pred_htmls.extend(self.table_matcher(structs, cells, dt_boxes, rec_res) if ocr_results is not None else [])
        res.cell_bboxes.extend(cells)
        res.logic_points.extend(self.table_matcher.decode_logic_points(structs))
        res.elapse = (time.perf_counter() - t0) / max(len(imgs), 1)
        return res


class B200RapidTableModel:
    """`RapidTableModel` for `model_type = slanet_1m` (rapid_doc/model/table/rapid_table.py:18-285): `predict(image RGB,
    ocr_result)` -> html.  Flow of :120-285 for this model type: OCR results (given, or from the injected `ocr_engine` and then
    text-normalised) -> white-out + placeholder boxes of embedded images (`fill_image_res`) -> formula / checkbox boxes
    (`mfd_res`) appended as OCR entries -> SLANet structure + matching (`B200RapidTable`) -> cell-text normalisation.
    The orientation pre-step (:131-162) runs when no OCR result is given and an `ocr_engine` is.  The wired / classifier branches
    need weights that are not available and are not built."""

    def __init__(self, ocr_engine=None, device=0, model_path=None, inline_delimiters=("$", "$")):
        self.ocr_engine = ocr_engine
        self.table_model = B200RapidTable(device=device, model_path=model_path, model_type="slanet_1m")
        self.inline_left, self.inline_right = inline_delimiters

    def batch_predict(self, images, ocr_result=None, fill_image_res=None, mfd_res=None, skip_text_in_image=True, use_img2table=False,
                      skip_table_orientation=None):
        return [self.predict(im, ocr_result, fill_image_res, mfd_res, skip_text_in_image, use_img2table, skip_table_orientation) for im in images]

    def predict(self, image, ocr_result=None, fill_image_res=None, mfd_res=None, skip_text_in_image=True, use_img2table=False,
                skip_table_orientation=None):
        import cv2
        from .table import needs_orientation_cls
        bgr = cv2.cvtColor(np.asarray(image), cv2.COLOR_RGB2BGR)
        if skip_table_orientation is None:
            skip_table_orientation = ocr_result is not None
        if not skip_table_orientation and self.ocr_engine is not None and bgr.shape[0] / max(bgr.shape[1], 1) > 1.2:
            det_res = self.ocr_engine.ocr(bgr, rec=False)[0]
            vertical = sum(1 for p1, _p2, p3, _p4 in (det_res or []) if ((p3[0] - p1[0]) / (p3[1] - p1[1]) if (p3[1] - p1[1]) > 0 else 1.0) < 0.8)
            if det_res and vertical >= len(det_res) * 0.3:                       # (this caller's threshold is 0.3, without the >= 3 rule)
                image = cv2.rotate(np.asarray(image), cv2.ROTATE_90_CLOCKWISE)
                bgr = cv2.cvtColor(image, cv2.COLOR_RGB2BGR)
        if not ocr_result:
            res = self.ocr_engine.ocr(bgr, mfd_res=mfd_res)[0] if self.ocr_engine is not None else None
            ocr_result = [list(x) for x in zip(*[[it[0], normalize_table_ocr_text(it[1][0]), it[1][1]] for it in res])] if res else None
        if not ocr_result:
            return None
        ocr_result = [list(ocr_result[0]), list(ocr_result[1]), list(ocr_result[2])]
        for fill in fill_image_res or []:
            bb = points_to_bbox(fill["ocr_bbox"])
            cv2.rectangle(bgr, (int(bb[0]), int(bb[1])), (int(bb[2]), int(bb[3])), (255, 255, 255), thickness=-1)
            ocr_result[0].append(fill["ocr_bbox"])
            ocr_result[1].append(fill["uuid"])
            ocr_result[2].append(1)
            if skip_text_in_image:
                inside = [i for i, o in enumerate(ocr_result[0][:-1])
                          if all(a >= b for a, b in zip(points_to_bbox(o)[:2], bb[:2])) and all(a <= b for a, b in zip(points_to_bbox(o)[2:], bb[2:]))]
                for i in sorted(inside, reverse=True):
                    for col in ocr_result:
                        del col[i]
        for mfd in mfd_res or []:
            if mfd.get("latex"):
                ocr_result[1].append(normalize_table_ocr_text(f"{self.inline_left}{mfd['latex']}{self.inline_right}"))
            elif mfd.get("checkbox"):
                ocr_result[1].append(normalize_table_ocr_text(mfd["checkbox"]))
            else:
                continue
            ocr_result[0].append(bbox_to_points(mfd["bbox"]))
            ocr_result[2].append(1)
        html_code = self.table_model([bgr], [ocr_result]).pred_htmls[0]
        return normalize_table_html_cell_text(html_code)
