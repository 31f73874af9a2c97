"""T7 / T1 (wireless tables): TableMatch restatement against the reference's own class (imported by path in the build container)
and against invariants that hold everywhere."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rapiddoc_b200 import table_match as TM       # noqa: E402

REF = "/root/reference/rapid_doc/model/table/rapid_table_self/table_matcher"


def _ref_matcher():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not mounted")
    import types
    pkg = types.ModuleType("ref_tm")
    pkg.__path__ = [REF]
    sys.modules["ref_tm"] = pkg
    for name in ("utils", "main"):
        spec = importlib.util.spec_from_file_location(f"ref_tm.{name}", os.path.join(REF, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"ref_tm.{name}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["ref_tm.main"].TableMatch()


def _case(seed, rows=5, cols=4, spans=False):
    rng = np.random.RandomState(seed)
    W, H = 600, 300
    xs, ys = np.linspace(10, W - 10, cols + 1), np.linspace(40, H - 10, rows + 1)
    tokens, cells = ["<html>", "<body>", "<table>", "<thead>"], []
    for r in range(rows):
        tokens.append("<tr>")
        c = 0
        while c < cols:
            span = 2 if spans and c + 1 < cols and rng.rand() < 0.25 else 1
            if span == 1:
                tokens.append("<td></td>")
            else:
                tokens += ["<td", f' colspan="{span}"', ">", "</td>"]
            cells.append([xs[c] + rng.randn(), ys[r] + rng.randn(), xs[c + span] + rng.randn(), ys[r + 1] + rng.randn()])
            c += span
        tokens.append("</tr>")
        if r == 0:
            tokens += ["</thead>", "<tbody>"]
    tokens += ["</tbody>", "</table>", "</body>", "</html>"]
    cells = np.array(cells)
    dt, rec = [], []
    for k, cb in enumerate(cells):
        n = rng.randint(0, 3)
        for j in range(n):
            w = (cb[2] - cb[0]) / max(n, 1)
            dt.append([cb[0] + j * w + 2, cb[1] + 3, cb[0] + (j + 1) * w - 2, cb[3] - 3])
            txt = f" t{k}_{j} " if rng.rand() < 0.3 else f"t{k}_{j}"
            if j == 0 and rng.rand() < 0.2:
                txt = "<b>" + txt + "</b>"
            rec.append((txt, float(rng.rand())))
    dt.append([5, 2, 100, 20]); rec.append(("title above the table", 0.9))          # filtered: ends above the first cell
    dt.append([W + 50, H + 50, W + 90, H + 70]); rec.append(("outside", 0.9))        # no intersection with any cell
    dt.append([xs[1] - 5, ys[1] - 4, xs[1] + 5, ys[1] + 4]); rec.append(("corner", 0.8))   # touches four cells: tie-breaks
    return (tokens, 0.9), cells, np.array(dt, np.float64), rec


@pytest.mark.parametrize("seed,spans", [(0, False), (1, True), (2, True), (3, False)])
def test_table_match_equals_the_reference(seed, spans):
    ref = _ref_matcher()
    struct, cells, dt, rec = _case(seed, spans=spans)
    mine = TM.TableMatch()
    assert mine([struct], [cells], [dt], [rec]) == ref([struct], [cells], [dt], [rec])
    f_dt, f_rec = mine.filter_ocr_result(cells, dt, rec)
    r_dt, r_rec = ref.filter_ocr_result(cells, dt, rec)
    assert np.array_equal(f_dt, r_dt) and f_rec == r_rec
    assert mine.match_result(cells, f_dt) == ref.match_result(cells, r_dt)
    assert np.array_equal(mine.decode_logic_points([struct])[0], ref.decode_logic_points([struct])[0])
    cells8 = np.stack([cells[:, 0], cells[:, 1], cells[:, 2], cells[:, 1], cells[:, 2], cells[:, 3], cells[:, 0], cells[:, 3]], 1)
    assert mine.match_result(cells8, f_dt) == ref.match_result(cells8, r_dt)
    assert mine([struct], [cells], [None], [None]) == [None]


def test_table_match_invariants():
    struct, cells, dt, rec = _case(5)
    m = TM.TableMatch()
    html = m.process_one(struct, cells, dt, rec)
    assert html.startswith("<html><body><table><tr><td>") and html.endswith("</table></body></html>")
    assert "<thead>" not in html and "title above the table" not in html and "outside" not in html
    assert html.count("<td>") == len(cells) and html.count("<tr>") == 5
    matched = m.match_result(cells, m.filter_ocr_result(cells, dt, rec)[0])
    assert sorted(i for v in matched.values() for i in v) == sorted(set(i for v in matched.values() for i in v))   # each box once
    lp = m.decode_one_logic_points(["<tr>", "<td", ' rowspan="2"', ">", "</td>", "<td></td>", "</tr>", "<tr>", "<td></td>", "</tr>"])
    assert lp == [[0, 1, 0, 0], [0, 0, 1, 1], [1, 1, 1, 1]]
    assert m.match_result(np.zeros((0, 4)), dt) == {} and m.match_result(cells, np.zeros((0, 4))) == {}


def test_format_ocr_results_clips_to_the_image():
    boxes = [[[-5, 3], [50, 3], [50, 20], [-5, 20]], [[90, 40], [130, 40], [130, 70], [90, 70]]]
    dt, rec = TM.format_ocr_results((boxes, ("a", "b"), (0.9, 0.8)), 60, 120)
    assert dt.tolist() == [[0, 3, 50, 20], [90, 40, 120, 60]] and rec == [("a", 0.9), ("b", 0.8)]


@pytest.mark.gpu
def test_rapid_table_end_to_end_on_synthetic_tables():
    """image + OCR results -> html through the CUDA structure model; every cell's text lands in its own <td>."""
    from rapiddoc_b200 import synth
    imgs = [synth.table_image(0, 4, 3), synth.table_image(1, 6, 4, 360, 520)]
    rt = TM.B200RapidTable(device=0)
    plain = rt(imgs)
    assert plain.pred_htmls == [] and [len(c) for c in plain.cell_bboxes] == [12, 24]
    assert plain.logic_points[0].tolist()[:4] == [[0, 0, 0, 0], [0, 0, 1, 1], [0, 0, 2, 2], [1, 1, 0, 0]]
    ocr = []
    for cells in plain.cell_bboxes:                   # one OCR box inside every predicted cell, text = its index
        boxes = [[[c[0] + 4, c[1] + 4], [c[2] - 4, c[1] + 4], [c[2] - 4, c[3] - 4], [c[0] + 4, c[3] - 4]] for c in cells]
        ocr.append((boxes, tuple(f"c{k}" for k in range(len(cells))), tuple(0.9 for _ in cells)))
    out = rt(imgs, ocr)
    for html, cells in zip(out.pred_htmls, out.cell_bboxes):
        got = [t.split("</td>")[0] for t in html.split("<td>")[1:]]
        assert got == [f"c{k}" for k in range(len(cells))]


def _ref_table_utils(monkeypatch):
    """rapid_doc/model/table/utils.py with its bs4 import satisfied by a stub (only the regex functions are used)."""
    path = "/root/reference/rapid_doc/model/table/utils.py"
    if not os.path.isfile(path):
        pytest.skip("reference tree not mounted")
    import types
    bs4 = types.ModuleType("bs4")
    bs4.BeautifulSoup = bs4.NavigableString = object
    monkeypatch.setitem(sys.modules, "bs4", bs4)
    spec = importlib.util.spec_from_file_location("ref_table_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_table_text_normalisation_equals_the_reference(monkeypatch):
    ref = _ref_table_utils(monkeypatch)
    texts = [None, 5, "", " 5號 ", "10號", "第6號", "香", "香 ", "哦樂", "a<b & 'c' \"d\"", "  plain  ", "号", "7號x"]
    for t in texts:
        assert TM.normalize_table_ocr_text(t) == ref.normalize_table_ocr_text(t), t
    cells = ["", None, "abc def", "中 文", "中 文 abc 测试 ， 好", "价格 $ 5 元", "（ 甲 ） 乙 ： 丙", "A 股 2024 年 报", "x ， y", "全角 ： 半角:", "中\t文\n换 行"]
    for c in cells:
        assert TM.normalize_table_cell_text(c) == ref.normalize_table_cell_text(c), c


def test_table_html_cell_normalisation():
    """The bs4-based upstream function cannot run here (parity unpinned): behaviour pinned by construction — only text nodes
    directly inside <td> / <th> change, attributes and untouched htmls come back identical."""
    h = '<html><body><table><tr><td>中 文 x</td><td colspan="2">a b</td><th>甲 ， 乙</th></tr><tr><td><b>粗 体</b></td><td>1 &lt; 2 中 文</td></tr></table></body></html>'
    out = TM.normalize_table_html_cell_text(h)
    assert out == '<html><body><table><tr><td>中文x</td><td colspan="2">a b</td><th>甲，乙</th></tr><tr><td><b>粗 体</b></td><td>1 &lt; 2中文</td></tr></table></body></html>'
    plain = "<html><body><table><tr><td>a b</td><td></td></tr></table></body></html>"
    assert TM.normalize_table_html_cell_text(plain) is plain and TM.normalize_table_html_cell_text("") == "" and TM.normalize_table_html_cell_text(None) is None
    assert TM.points_to_bbox([[1, 2], [9, 2], [9, 7], [1, 7]]) == [1, 2, 9, 7] and TM.bbox_to_points([1, 2, 9, 7]).tolist() == [[1, 2], [9, 2], [9, 7], [1, 7]]


@pytest.mark.gpu
def test_rapid_table_model_predict_on_a_synthetic_table():
    """RapidTableModel.predict for slanet_1m: RGB crop + OCR results (+ a formula box, + an embedded image) -> html."""
    import cv2
    from rapiddoc_b200 import synth
    img = synth.table_image(0, 4, 3)
    rgb = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    m = TM.B200RapidTableModel(device=0)
    cells = m.table_model([img]).cell_bboxes[0]
    boxes = [[[c[0] + 4, c[1] + 4], [c[2] - 4, c[1] + 4], [c[2] - 4, c[3] - 4], [c[0] + 4, c[3] - 4]] for c in cells]
    texts = [f"单 元 {k}" if k % 2 else f"c{k} <x>" for k in range(len(cells))]
    ocr = [boxes[:-2], texts[:-2], [0.9] * (len(cells) - 2)]
    mfd = [{"bbox": [cells[-2][0] + 4, cells[-2][1] + 4, cells[-2][2] - 4, cells[-2][3] - 4], "latex": "a^2"}]
    fill = [{"ocr_bbox": boxes[-1], "uuid": "img-uuid-1"}]
    html = m.predict(rgb, ocr_result=ocr, fill_image_res=fill, mfd_res=mfd)
    got = [t.split("</td>")[0] for t in html.split("<td>")[1:]]
    assert got[0] == "c0 <x>" and got[1] == "单元1" and got[-2] == "$a^2$" and got[-1] == "img-uuid-1" and len(got) == 12
    assert m.predict(rgb, ocr_result=None) is None                                    # no OCR engine injected and no result given
    assert m.batch_predict([rgb], ocr_result=ocr)[0].count("<td>") == 12


@pytest.mark.gpu
def test_table_custom_model_plugin_runs_ocr_and_structure():
    """The CustomBaseModel table plugin: RGB table crops in, html with the recognised cell texts out (OCR + SLANet on the GPU)."""
    import cv2
    from rapiddoc_b200 import synth
    from rapiddoc_b200.plugin import B200TableCustomModel, CustomBaseModel
    m = B200TableCustomModel(device=0)
    assert isinstance(m, CustomBaseModel)
    imgs = [cv2.cvtColor(synth.table_image(s, 4, 3, 360, 560), cv2.COLOR_BGR2RGB) for s in (0, 1)]
    htmls = m.batch_predict(imgs, fill_image_res_list=[[], []])
    for h in htmls:
        assert h.startswith("<html><body><table>") and h.count("<tr>") == 4 and h.count("<td>") == 12
        filled = [t.split("</td>")[0] for t in h.split("<td>")[1:]]
        assert sum(1 for t in filled if t.strip()) >= 9                      # the synthetic cell texts were recognised and matched


def test_match_result_fuzz_against_the_reference():
    """Random cell grids and OCR boxes (overlapping, touching, degenerate, far away): same assignment as the reference."""
    ref = _ref_matcher()
    rng = np.random.RandomState(7)
    mine = TM.TableMatch()
    for trial in range(150):
        k, m = rng.randint(1, 12), rng.randint(0, 15)
        x0, y0 = rng.randint(0, 300, k), rng.randint(0, 200, k)
        cells = np.stack([x0, y0, x0 + rng.randint(0, 120, k), y0 + rng.randint(0, 60, k)], 1).astype(np.float64)
        if trial % 3 == 0:
            cells[rng.randint(0, k)] = cells[rng.randint(0, k)]                      # duplicate cells: index tie-break
        bx, by = rng.randint(-20, 320, m), rng.randint(-20, 220, m)
        dt = np.stack([bx, by, bx + rng.randint(0, 90, m), by + rng.randint(0, 40, m)], 1).astype(np.float64)
        if trial % 4 == 0 and m:
            dt[0] = cells[0]                                                          # an exact hit
        assert mine.match_result(cells, dt) == ref.match_result(cells, dt), (cells, dt)
