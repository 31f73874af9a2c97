"""Table cell / OCR matching and HTML assembly (wireless-table half) and the SLANet orchestration of `RapidTable`
(SURVEY T1), restated so that the table path runs without an importable `rapid_doc`:

  TableMatch           rapid_table_self/table_matcher/main.py:120-348  (filter_ocr_result, match_result with the
                       (1 - IoU, distance) ordering of :75-117, get_pred_html, decode_logic_points)
  format_ocr_results   rapid_table_self/utils/utils.py:15-27
  B200RapidTable       RapidTable.__call__ for the PP-structure models, rapid_table_self/main.py:78-123

Host glue (a few hundred boxes per table); the structure model under it is `table.B200TableStructurer` (CUDA).
tests/test_table_match.py runs it against the reference's own class imported by path.
"""
import time

import numpy as np

MIN_IOU = 0.1 ** 8


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
