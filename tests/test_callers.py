"""D0 / R0: rapiddoc_b200/window.py (restated `_run_ocr_det_batch` / `_run_ocr_rec_postprocess`) against the reference's REAL
functions (analyze_utils.py imported by path with stub parents, build container only) driven by a deterministic stand-in
model; the plugin install hooks against stub `rapid_doc` / `rapidocr` modules; and (GPU) the same callers driving
`B200OcrModel`, compared with the fused `ocr_pages` path and the CPU oracle."""
import copy
import os
import sys
import types

import numpy as np
import pytest

from rapiddoc_b200 import window

REF = "/root/reference"
HAVE_REF = os.path.exists(REF)


class FakeOcr:
    """Deterministic stand-in with the RapidOcrModel surface the callers use."""

    def __init__(self):
        self.det_calls, self.rec_calls = [], []

    def det_batch_predict(self, imgs, max_batch_size=8):
        self.det_calls.append((len(imgs), imgs[0].shape, max_batch_size))
        out = []
        for im in imgs:
            h, w = im.shape[:2]
            s = int(im[::7, ::7].astype(np.int64).sum() % 5)
            boxes = [np.array([[60 + s, 55 + 30 * k], [w - 70, 55 + 30 * k + (2 if k == 1 else 0)], [w - 70, 78 + 30 * k + (2 if k == 1 else 0)], [60 + s, 78 + 30 * k]], np.float32)
                     for k in range(max(1, (h - 110) // 30))]
            if s == 3:
                boxes.append(np.array([[80, 60], [110, 50], [130, 120], [100, 130]], np.float32))   # a rotated one
            out.append((np.array(boxes[::-1]), 0.01) if s != 4 else (None, 0))
        return out

    def ocr(self, imgs, det=False, rec=True, tqdm_enable=False, **kw):
        assert det is False and isinstance(imgs, list)
        self.rec_calls.append(len(imgs))
        res = []
        for im in imgs:
            h, w = im.shape[:2]
            res.append((f"t{h}x{w}" if w % 3 else "号", 0.3 + (int(im.sum()) % 70) / 100.0))
        return [res]


def pages(seed=0, n=3):
    rng = np.random.default_rng(seed)
    out = []
    for p in range(n):
        img = rng.integers(0, 256, (700, 900, 3), dtype=np.uint8)
        blocks = []
        y = 20
        for b in range(int(rng.integers(2, 5))):
            h, w = int(rng.integers(60, 200)), int(rng.integers(200, 700))
            x = int(rng.integers(0, 150))
            blocks.append({"poly": [x, y, x + w, y, x + w, y + h, x, y + h], "original_label": "text", "original_order": b,
                           **({"need_ocr_det": True} if (p == 1 and b % 2 == 0) else {})})
            y += h + 10
            if y > 480:
                break
        out.append({"ocr_res_list": blocks, "ocr_enable": p != 1, "np_img": img, "lang": "ch" if p != 2 else "en", "layout_res": [],
                    "single_page_mfdetrec_res": [{"bbox": [100, 40, 180, 70]}] if p == 0 else [], "checkbox_res": []})
    return out


def _strip(layout_res):
    return [{k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in d.items()} for d in layout_res]


def _reference_analyze_utils():
    def stub(name, path=None):
        m = types.ModuleType(name)
        if path:
            m.__path__ = [path]
        sys.modules[name] = m
        return m
    saved = {k: v for k, v in sys.modules.items() if k == "rapid_doc" or k.startswith("rapid_doc.")}
    for k in saved:
        del sys.modules[k]
    stub("rapid_doc", f"{REF}/rapid_doc")
    for sub in ["backend", "backend.pipeline", "utils", "model", "model.table"]:
        stub("rapid_doc." + sub, f"{REF}/rapid_doc/" + sub.replace(".", "/"))
    mi = stub("rapid_doc.backend.pipeline.model_init")

    class AtomModelSingleton:
        model = None

        def get_atom_model(self, **kw):
            return AtomModelSingleton.model
    mi.AtomModelSingleton = AtomModelSingleton
    stub("rapid_doc.model.table.utils").normalize_table_ocr_text = lambda x: x
    sp = stub("rapid_doc.utils.span_pre_proc")
    for k in ["txt_spans_extract", "txt_spans_bbox_extract", "txt_most_angle_extract_table", "extract_table_fill_image"]:
        setattr(sp, k, None)
    import rapid_doc.backend.pipeline.analyze_utils as au
    return au, AtomModelSingleton, saved


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_window_callers_equal_reference_callers():
    au, Singleton, saved = _reference_analyze_utils()
    try:
        for cfg in ({}, {"use_det_mode": "ocr", "Det.rec_batch_num": 4}, {"use_det_mode": "txt"}):
            a, b = pages(), pages()
            fa, fb = FakeOcr(), FakeOcr()
            Singleton.model = fa
            au._run_ocr_det_batch(a, Singleton(), cfg)
            window.run_ocr_det_batch(b, lambda lang: fb, cfg)
            assert fa.det_calls == fb.det_calls
            for pa, pb in zip(a, b):
                assert _strip(pa["layout_res"]) == _strip(pb["layout_res"])
            la, lb = [p["layout_res"] for p in a], [p["layout_res"] for p in b]
            au._run_ocr_rec_postprocess(la, cfg)
            window.run_ocr_rec_postprocess(lb, lambda lang: fb, cfg)
            assert fa.rec_calls == fb.rec_calls and [_strip(x) for x in la] == [_strip(x) for x in lb]
            if cfg.get("use_det_mode") != "txt":
                assert sum(len(x) for x in la) > 5
    finally:
        for k in [k for k in sys.modules if k == "rapid_doc" or k.startswith("rapid_doc.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_plugin_install_hooks_against_stub_modules(monkeypatch):
    """install() / install_window() / install_engine() rebind exactly the names the reference looks up."""
    from rapiddoc_b200 import B200Error, _lib, plugin
    made = []

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    class Singleton:
        _models = {"cached": 1}

        def get_atom_model(self, **kw):
            made.append(kw)
            return FakeOcr()
    orig_init = lambda *a, **k: "reference-model"          # noqa: E731
    stub("rapid_doc"), stub("rapid_doc.backend")
    mi = stub("rapid_doc.backend.pipeline.model_init", ocr_model_init=orig_init, AtomModelSingleton=Singleton)
    ml = stub("rapid_doc.backend.pipeline.model_list", AtomicModel=types.SimpleNamespace(OCR="ocr"))
    au = stub("rapid_doc.backend.pipeline.analyze_utils", _run_ocr_det_batch="d0", _run_ocr_rec_postprocess="r0", AtomModelSingleton=Singleton)
    ba = stub("rapid_doc.backend.pipeline.batch_analyze", _run_ocr_det_batch="d0", _run_ocr_rec_postprocess="r0")
    stub("rapid_doc.backend.pipeline", model_init=mi, model_list=ml, analyze_utils=au, batch_analyze=ba)
    assert plugin.install(device=0) is orig_init
    assert mi.ocr_model_init is not orig_init and Singleton._models == {}
    assert mi.ocr_model_init(is_seal=True) == "reference-model"            # seal OCR keeps the reference implementation
    if _lib.load().rdb_device_count() == 0:
        with pytest.raises(B200Error):                                       # no silent CPU fallback
            mi.ocr_model_init(det_db_box_thresh=0.3)
    assert plugin.install_window() == ("d0", "r0")
    assert callable(au._run_ocr_det_batch) and ba._run_ocr_det_batch is au._run_ocr_det_batch
    pg = pages(1, 2)
    au._run_ocr_det_batch(pg, Singleton(), {})
    au._run_ocr_rec_postprocess([p["layout_res"] for p in pg], {})
    assert made and made[0]["atom_model_name"] == "ocr" and all("text" in d for p in pg for d in p["layout_res"])
    # engine seam: rapidocr's torch session class is replaced; det / rec picked by the model file name
    rt = stub("rapidocr.inference_engine.pytorch.main", TorchInferSession=object)
    stub("rapidocr"), stub("rapidocr.inference_engine")
    stub("rapidocr.inference_engine.pytorch", main=rt, TorchInferSession=object)
    plugin.install_engine()
    assert rt.TorchInferSession is not object
    if _lib.load().rdb_device_count() == 0:
        with pytest.raises(B200Error):
            rt.TorchInferSession({"model_path": os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights", "ch_PP-OCRv6_rec_small.safetensors")})


def test_plugin_table_and_orientation_hooks_against_stub_modules(monkeypatch):
    from rapiddoc_b200 import B200Error, _lib, plugin

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    class Singleton:
        _models = {"cached": 1}

    class RapidTable:
        def __init__(self, model_type, path=None):
            self.cfg = types.SimpleNamespace(model_type=types.SimpleNamespace(value=model_type), model_dir_or_path=path)
            self.table_structure = self._init_table_structer()

        def _init_table_structer(self):
            return "reference-structurer"
    stub("rapid_doc"), stub("rapid_doc.backend"), stub("rapid_doc.model"), stub("rapid_doc.model.table")
    mi = stub("rapid_doc.backend.pipeline.model_init", img_orientation_cls_model_init=lambda: "reference-orientation", AtomModelSingleton=Singleton)
    stub("rapid_doc.backend.pipeline", model_init=mi)
    rt = stub("rapid_doc.model.table.rapid_table_self.main", RapidTable=RapidTable)
    stub("rapid_doc.model.table.rapid_table_self", main=rt)
    assert plugin.install_orientation(device=0)() == "reference-orientation" and Singleton._models == {}
    plugin.install_table_structure(device=0)
    assert RapidTable("unet").table_structure == "reference-structurer"          # other model types keep the reference object
    if _lib.load().rdb_device_count() == 0:
        with pytest.raises(B200Error):                                            # no silent CPU fallback
            mi.img_orientation_cls_model_init()
        with pytest.raises(B200Error):
            RapidTable("slanet_1m", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights", "slanet-1m.onnx"))
    else:
        assert type(RapidTable("slanet_1m").table_structure).__name__ == "B200TableStructurer"
        assert type(mi.img_orientation_cls_model_init()).__name__ == "B200OrientationModel"


def test_multi_gpu_dispatch_with_stand_in_models():
    """B200OcrPool: pages round-robin, whole reference rec batches dealt round-robin, results in input order."""
    from rapiddoc_b200.multi import B200OcrPool

    class M(FakeOcr):
        def __init__(self, dev):
            super().__init__()
            self.dev, self.rec_batch_num, self.drop_score = dev, 4, 0.5

        def ocr_pages(self, pg, mfd=None, drop=None):
            return [[(self.dev, p.shape)] for p in pg]
    pool = B200OcrPool([0, 1, 2], model_factory=M)
    imgs = [np.full((120 + 10 * i, 300, 3), i, np.uint8) for i in range(10)]
    one = FakeOcr().det_batch_predict(imgs)
    got = pool.det_batch_predict(imgs)
    assert all((a[0] is None and b[0] is None) or np.array_equal(a[0], b[0]) for a, b in zip(got, one))
    assert [m.det_calls[0][0] for m in pool.models] == [4, 3, 3]
    crops = [np.full((20, 40 + 13 * ((7 * i) % 23), 3), i, np.uint8) for i in range(23)]
    want = FakeOcr().ocr(crops, det=False)[0]
    assert pool.ocr(crops, det=False)[0] == want
    shares = pool.rec_shares([c.shape[:2] for c in crops])
    order = np.argsort(np.array([c.shape[1] / c.shape[0] for c in crops]))
    assert sorted(sum(shares, [])) == list(range(23)) and shares[0][:4] == [int(i) for i in order[:4]] and shares[1][:4] == [int(i) for i in order[4:8]]
    assert [r[0][0] for r in pool.ocr_pages(imgs)] == [i % 3 for i in range(10)]
    pool.close()


@pytest.mark.gpu
def test_gpu_callers_drive_b200_model():
    """The restated callers on top of B200OcrModel (fp32): every detected line is recognised; texts equal the CPU oracle's on
    the same crops (the callers crop on the host exactly as the reference does)."""
    from oracle import pipeline as OP
    from rapiddoc_b200 import PREC_FP32, synth
    from rapiddoc_b200.ocr import B200OcrModel
    page = synth.det_pages(1, 512, 768, seed=9, lines=10)[0]
    rgb = page[:, :, ::-1].copy()
    info = [{"ocr_res_list": [{"poly": [0, 0, 768, 0, 768, 250, 0, 250], "original_label": "text", "original_order": 0},
                              {"poly": [0, 250, 768, 250, 768, 512, 0, 512], "original_label": "text", "original_order": 1}],
             "ocr_enable": True, "np_img": rgb, "lang": "ch", "layout_res": [], "single_page_mfdetrec_res": [], "checkbox_res": []}]
    model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, precision=PREC_FP32)
    window.run_ocr_det_batch(info, lambda lang: model, {})
    spans = info[0]["layout_res"]
    assert len(spans) >= 6 and all(s["category_id"] == 15 and "np_img" in s for s in spans)
    crops = [s["np_img"] for s in spans]
    want = OP.rec_crops(crops, rec_batch_num=6)
    window.run_ocr_rec_postprocess([spans], lambda lang: model, {})
    assert [s["text"] for s in spans] == [t for t, _ in want]
    assert all(abs(s["score"] - float(f"{c:.3f}")) <= 1e-3 for s, (_, c) in zip(spans, want))
