"""Layout (L1/L2/L4/L5) and table (T2/T3/T5/T8) numerics against the reference's own functions.

The reference modules are imported by path from /root/reference when that tree is mounted (build container): CPU tests compare
directly, and `python tests/test_layout_table.py` writes tests/golden/layout_table_cases.npz from them.  The GPU tests (and the CPU
tests on a box without the reference) compare against that committed fixture."""
import importlib.util
import os
import pickle
import sys
import types

import numpy as np
import pytest

from rapiddoc_b200 import layout as L
from rapiddoc_b200 import table as T

REF = "/root/reference/rapid_doc"
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layout_table_cases.npz")
HAVE_REF = os.path.exists(REF)

LABELS_V2 = ["abstract", "algorithm", "aside_text", "chart", "content", "display_formula", "doc_title", "figure_title", "footer", "footer_image",
             "footnote", "formula_number", "header", "header_image", "image", "inline_formula", "number", "paragraph_title", "reference",
             "reference_content", "seal", "table", "text", "vertical_text", "vision_footnote"]
MERGE_V2 = {i: "union" for i in range(25)}
MERGE_V2.update({3: "large", 5: "large", 6: "large", 15: "large", 17: "large"})
STRUCT_VOCAB = ["<thead>", "</thead>", "<tbody>", "</tbody>", "<tr>", "</tr>", "<td>", "<td", ">", "</td>"] + [f' colspan="{i}"' for i in range(2, 21)] + \
               [f' rowspan="{i}"' for i in range(2, 21)]


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def ref_post():
    return _load(f"{REF}/model/layout/rapid_layout_self/model_handler/pp_doclayout/post_process.py", "ref_layout_post")


def ref_table_post():
    # post_process.py imports `...utils.typings.ModelType` and `..utils.wrap_with_html_struct` relatively: give it stub parents
    for name in ["rt", "rt.utils", "rt.table_structure", "rt.table_structure.pp_structure"]:
        sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules[name].__path__ = []
    ty = types.ModuleType("rt.utils.typings")

    class ModelType:
        SLANETPLUS = "slanet_plus"
    ty.ModelType = ModelType
    sys.modules["rt.utils.typings"] = ty
    ut = types.ModuleType("rt.table_structure.utils")
    ut.wrap_with_html_struct = lambda s: ["<html>", "<body>", "<table>"] + s + ["</table>", "</body>", "</html>"]
    sys.modules["rt.table_structure.utils"] = ut
    spec = importlib.util.spec_from_file_location("rt.table_structure.pp_structure.post_process",
                                                  f"{REF}/model/table/rapid_table_self/table_structure/pp_structure/post_process.py")
    m = importlib.util.module_from_spec(spec)
    m.__package__ = "rt.table_structure.pp_structure"
    spec.loader.exec_module(m)
    return m, ModelType


def box_sets(seed=0, pages=12):
    """Seeded detector outputs [n,7] float32 (cls, score, x1, y1, x2, y2, order): clusters of overlapping boxes so that the same-class
    0.6 rule, the cross-class 0.98 rule and containment all fire; duplicated scores exercise the tie order."""
    rng = np.random.default_rng(seed)
    out = []
    for p in range(pages):
        rows = []
        for _ in range(int(rng.integers(3, 40))):
            x, y = rng.uniform(0, 1400), rng.uniform(0, 2000)
            w, h = rng.uniform(20, 600), rng.uniform(10, 300)
            cls = int(rng.integers(0, 25))
            for k in range(int(rng.integers(1, 5))):
                jx, jy, jw, jh = rng.normal(0, 6), rng.normal(0, 6), rng.normal(0, 8), rng.normal(0, 8)
                c = cls if rng.random() < 0.7 else int(rng.integers(0, 25))
                sc = float(rng.choice([0.31, 0.5, 0.75])) if rng.random() < 0.15 else rng.uniform(0.05, 0.99)
                shrink = rng.uniform(0.3, 1.0) if rng.random() < 0.3 else 1.0
                rows.append([c, sc, x + jx, y + jy, x + jx + (w + jw) * shrink, y + jy + (h + jh) * shrink, rng.integers(0, 1000)])
        if p == 3:
            rows.append([14, 0.9, -5, -5, 1700, 2400, 7])          # a page-sized `image` box
        out.append(np.array(rows, np.float32))
    return out


IMG_SIZES = [(1654, 2339)] * 6 + [(2339, 1654)] * 6


def reference_cases():
    R = ref_post()
    sets = box_sets()
    out = {"n": len(sets)}
    for p, b in enumerate(sets):
        thr = b[(b[:, 1] > 0.3) & (b[:, 0] > -1)]
        out[f"nms{p}"] = np.array(R.nms(thr[:, :6].copy(), iou_same=0.6, iou_diff=0.98), np.int64)
        co, cb = R.check_containment(thr[:, :6], None)
        out[f"co{p}"], out[f"cb{p}"] = co, cb
        co, cb = R.check_containment(thr[:, :6], 5, 17, "large")
        out[f"col{p}"], out[f"cbl{p}"] = co, cb
        co, cb = R.check_containment(thr[:, :6], 5, 3, "small")
        out[f"cos{p}"], out[f"cbs{p}"] = co, cb
        pp = R.PPPostProcess(LABELS_V2, 0.3, 0.5, layout_merge_bboxes_mode=MERGE_V2, layout_unclip_ratio=[1.0, 1.0], scale_size=(800, 800))
        res = pp(b.copy(), IMG_SIZES[p], None, "rect")
        out[f"post{p}"] = np.frombuffer(pickle.dumps([(d["cls_id"], d["label"], d["score"], d["coordinate"], d["order"]) for d in res]), np.uint8)
    # table decode
    TP, MT = ref_table_post()
    rng = np.random.default_rng(5)
    probs = rng.random((4, 60, 50)).astype(np.float32)
    probs[:, :, 49] *= 0.3
    probs[0, 20, 49] = 5.0
    probs[2, 7, 49] = 5.0
    td_empty, td_open = 48, 7            # "<td></td>" (appended last, before eos) and "<td" after sos + the merged vocabulary
    for b in range(4):
        probs[b, [1, 3, 5], td_empty] = 4.0
        probs[b, 2, td_open] = 4.0
    bbox = rng.random((4, 60, 8)).astype(np.float32)
    shapes = np.array([[300, 400, 1.2, 1.2, 488, 488]] * 4, np.float64)
    imgs = [np.zeros((300, 400, 3), np.uint8)] * 4
    dec = TP.TableLabelDecode(list(STRUCT_VOCAB), {"model_type": MT.SLANETPLUS})
    structs, cells = dec(bbox.copy(), probs.copy(), shapes, imgs)
    out["t_probs"], out["t_bbox"] = probs, bbox
    out["t_structs"] = np.frombuffer(pickle.dumps(structs), np.uint8)
    for i, c in enumerate(cells):
        out[f"t_cells{i}"] = c
    return out


def _fixture():
    return np.load(FIX, allow_pickle=False)


def _cases():
    return reference_cases() if HAVE_REF else _fixture()


# --------------------------------------------------------------------------------------- CPU
@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_fixture_matches_live_reference():
    live, fix = reference_cases(), _fixture()
    for k in fix.files:
        assert np.array_equal(np.asarray(live[k]), fix[k]), k


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_host_numerics_against_reference_functions():
    import cv2
    R = ref_post()
    # L2 preprocessing: import the class with a stub for its relative `..utils` import
    src = open(f"{REF}/model/layout/rapid_layout_self/model_handler/pp_doclayout/pre_process.py").read().replace("from ..utils import ModelType", "")
    ns = {}

    class MT:
        PP_DOCLAYOUT_L, PP_DOCLAYOUT_PLUS_L, PP_DOCLAYOUTV2, PP_DOCLAYOUTV3, PP_DOCLAYOUT_S = "l", "pl", "v2", "v3", "s"
    ns["ModelType"] = MT
    exec(compile(src, "ref_pre", "exec"), ns)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (517, 389, 3), dtype=np.uint8)
    for mt, size, inorm in ((MT.PP_DOCLAYOUTV3, (800, 800), False), (MT.PP_DOCLAYOUT_S, (480, 480), True)):
        want = ns["PPPreProcess"](size, mt)(img)
        got = L.LayoutPreProcess(size, imagenet_norm=inorm)(img)
        assert got.dtype == want.dtype and np.array_equal(got, want)
    # unclip_boxes / restructured_boxes
    b = box_sets(1, 2)[0][:, :6]
    assert np.array_equal(L.unclip_boxes(b, (1.3, 0.9)), R.unclip_boxes(b, (1.3, 0.9)))
    assert np.array_equal(L.unclip_boxes(b, {3.0: (1.5, 1.5)}), R.unclip_boxes(b, {3.0: (1.5, 1.5)}))
    want = R.restructured_boxes(b, LABELS_V2, (1000, 1500))
    assert L.restructured_boxes(b, LABELS_V2, (1000, 1500)) == want
    # T3 / T2 preprocessing
    tp = _load(f"{REF}/model/table/rapid_table_self/table_structure/pp_structure/pre_process.py", "ref_table_pre")
    imgs = [rng.integers(0, 256, (300, 411, 3), dtype=np.uint8), rng.integers(0, 256, (620, 233, 3), dtype=np.uint8)]
    wi, ws = tp.TablePreprocess()(imgs)
    gi, gs = T.TablePreprocess()(imgs)
    assert np.array_equal(ws, gs) and all(np.array_equal(a, b) for a, b in zip(wi, gi))


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_filter_overlap_boxes_against_reference():
    # backend/utils/utils.py imports heavy siblings at module level: exec only the function with its two helpers bound
    src = open(f"{REF}/backend/utils/utils.py").read()
    a = src.index("def filter_overlap_boxes(")
    b = src.index("def _rect_from_poly(")
    ru = _load(f"{REF}/model/reading_order/utils.py", "ref_ro_utils")
    R = ref_post()
    from copy import deepcopy
    ns = {"deepcopy": deepcopy, "calculate_overlap_ratio": ru.calculate_overlap_ratio, "calculate_bbox_area": R.calculate_bbox_area,
          "calculate_polygon_overlap_ratio": R.calculate_polygon_overlap_ratio}
    exec(compile(src[a:b], "ref_filter", "exec"), ns)
    rng = np.random.default_rng(2)
    for p, bs in enumerate(box_sets(3, 8)):
        res = []
        for r in bs[:40]:
            x1, y1, x2, y2 = [float(v) for v in r[2:6]]
            label = ["text", "image", "inline_formula", "reference", "table", "chart"][int(r[0]) % 6]
            res.append({"original_label": label, "poly": [x1, y1, x2, y1, x2, y2, x1, y2], "polygon_points": None, "score": float(r[1])})
        for custom in (False, True):
            assert L.filter_overlap_boxes(res, custom) == ns["filter_overlap_boxes"](res, custom)


def test_table_helpers_and_orientation_rule():
    lab, sc = T.cls_scores(np.array([[2.0, 1.0], [0.1, 0.3]], np.float32))
    assert lab == ["wired", "wireless"] and abs(sc[0] - 0.7310586) < 1e-6
    assert T.cls_vote(["wired", "wired"], [0.9, 0.8], ["wired", "wireless"], [0.7, 0.95]) == (["wired", "wireless"], [0.7, 0.8])
    tall = [[[0, 0], [10, 0], [10, 40], [0, 40]]] * 3 + [[[0, 0], [50, 0], [50, 10], [0, 10]]] * 5
    assert T.needs_orientation_cls((600, 400), tall) and not T.needs_orientation_cls((400, 600), tall)
    assert not T.needs_orientation_cls((600, 400), tall[3:]) and T.needs_orientation_cls((600, 400), None)
    m = L.class_map(LABELS_V2)
    assert m["text"] == 1 and m["header"] == 2 and m["table"] == 5 and m["display_formula"] == 14 and m["inline_formula"] == 13


def test_layout_model_refuses_without_weights():
    from rapiddoc_b200 import B200Error
    with pytest.raises(B200Error, match="weights unavailable"):
        L.B200LayoutModel(session=None, labels=LABELS_V2)


# --------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_nms_and_containment_bit_identical():
    g = _cases()
    sets = box_sets()
    thr = [b[(b[:, 1] > 0.3) & (b[:, 0] > -1)] for b in sets]
    kept = L.nms_window([t[:, :6] for t in thr], 0.6, 0.98)
    for p in range(len(sets)):
        assert kept[p] == np.asarray(g[f"nms{p}"]).tolist(), p
    for key, args in (("", (None, None, None)), ("l", (5, 17, "large")), ("s", (5, 3, "small"))):
        rel = L.containment_window([t[:, :6] for t in thr], *args)
        for p in range(len(sets)):
            assert np.array_equal(rel[p][0], g[f"co{key}{p}"]) and np.array_equal(rel[p][1], g[f"cb{key}{p}"]), (key, p)


@pytest.mark.gpu
def test_gpu_layout_postprocess_and_facade():
    g = _cases()
    sets = box_sets()
    post = L.LayoutPostProcess(LABELS_V2, 0.3, True, MERGE_V2, [1.0, 1.0], (800, 800))
    got = post([b.copy() for b in sets], IMG_SIZES)
    for p in range(len(sets)):
        want = pickle.loads(np.asarray(g[f"post{p}"]).tobytes())
        assert [(d["cls_id"], d["label"], d["score"], d["coordinate"], d["order"]) for d in got[p]] == want, p

    # facade around a stand-in session that returns the seeded boxes (the reference's InferSession contract)
    class Session:
        def __call__(self, x, sf):
            assert x.shape == (2, 3, 800, 800) and x.dtype == np.float32 and sf.shape == (2, 2)
            return [np.concatenate(sets[:2]), np.array([len(sets[0]), len(sets[1])])]
    model = L.B200LayoutModel(Session(), "pp_doclayoutv3", LABELS_V2, 0.3, MERGE_V2)
    pages = model.batch_predict([np.zeros((2339, 1654, 3), np.uint8)] * 2, batch_size=2)
    assert [len(p) for p in pages] == [len(got[0]), len(got[1])]
    assert pages[0][0]["original_order"] == 0 and set(pages[0][0]) == {"category_id", "original_label", "original_order", "poly", "polygon_points", "score"}


@pytest.mark.gpu
def test_gpu_table_label_decode():
    g = _cases()
    probs, bbox = np.asarray(g["t_probs"]), np.asarray(g["t_bbox"])
    dec = T.TableLabelDecode(list(STRUCT_VOCAB), slanet_plus=True)
    shapes = np.array([[300, 400, 1.2, 1.2, 488, 488]] * 4, np.float64)
    imgs = [np.zeros((300, 400, 3), np.uint8)] * 4
    idx, val = dec.argmax(probs)
    assert np.array_equal(idx, probs.argmax(2)) and np.array_equal(val, probs.max(2))
    structs, cells = dec(bbox.copy(), probs, shapes, imgs)
    assert structs == pickle.loads(np.asarray(g["t_structs"]).tobytes())
    for i, c in enumerate(cells):
        assert np.array_equal(c, g[f"t_cells{i}"])
    import torch
    structs2, _ = dec(bbox.copy(), torch.from_numpy(probs).cuda(), shapes, imgs)
    assert structs2 == structs


if __name__ == "__main__":
    np.savez_compressed(FIX, **{k: np.asarray(v) for k, v in reference_cases().items()})
    print("wrote", FIX)
