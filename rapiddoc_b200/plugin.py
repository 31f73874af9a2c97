"""Drop-in hooks: put the B200 engines under RapidDoc's unchanged orchestration.

Three nested seams, all provided here:

1. `CustomBaseModel.batch_predict(image_list, **kw) -> list[str]`
   (rapid_doc/model/custom/__init__.py:4-20; dispatched by
   rapid_doc/backend/pipeline/model_init.py:96-120 on isinstance).  `B200OcrCustomModel`
   is the block-level OCR plugin: each layout block image -> its text lines joined by "\n"
   (the mode `_run_custom_ocr` expects, rapid_doc/backend/pipeline/batch_analyze.py:286-333).
2. The `rapid_doc/model/*` class API: `install()` rebinds
   `rapid_doc.backend.pipeline.model_init.ocr_model_init` (model_init.py:45-54) so that
   AtomModelSingleton builds `B200OcrModel` instead of `RapidOcrModel` (seal models keep the
   reference implementation).  Nothing in the reference tree is edited.
3. The engine protocol `InferSession.__call__(np.ndarray) -> np.ndarray`:
   `install_engine()` swaps rapidocr's torch session class the same way
   rapid_doc/model/ocr/ocr_patch.py:95-105 does, with `B200DetSession` / `B200RecSession`
   picked by the model file name.
"""
from abc import ABC, abstractmethod

import numpy as np

try:  # the reference's ABC when RapidDoc is installed, an identical stand-in otherwise
    from rapid_doc.model.custom import CustomBaseModel
except Exception:  # pragma: no cover - exercised on boxes without RapidDoc
    class CustomBaseModel(ABC):
        @abstractmethod
        def batch_predict(self, image_list, **kwargs):
            ...


class B200OcrCustomModel(CustomBaseModel):
    """Block-level OCR plugin (seam 1): BGR block images -> multi-line text."""

    def __init__(self, device=0, precision=None, box_thresh=0.3, unclip_ratio=1.8):
        from .ocr import B200OcrModel
        self.model = B200OcrModel(det_db_box_thresh=box_thresh, det_db_unclip_ratio=unclip_ratio, device=device, precision=precision)

    def batch_predict(self, image_list, **kwargs):
        out = []
        for img in image_list:
            boxes, res = self.model(np.ascontiguousarray(img))
            out.append("\n".join(t for t, _ in res) if res else "")
        return out


class B200TableCustomModel(CustomBaseModel):
    """Table plugin (seam 1): `table_config={"custom_model": B200TableCustomModel(device=0)}` — `BatchAnalyze._run_table_recognition`
    then calls `batch_predict(table_imgs, fill_image_res_list=...)` (rapid_doc/backend/pipeline/batch_analyze.py:359-380) and stores
    the returned html.  Inside: the OCR model the reference's `table_model_init` builds for tables (box_thresh 0.5, unclip 1.6, no
    box merging; model_init.py:14-29) and `B200RapidTableModel` (SLANet-1m structure on the GPU + matching)."""

    def __init__(self, device=0, precision=None, ocr_model=None):
        from .table_match import B200RapidTableModel
        if ocr_model is None:
            from .ocr import B200OcrModel
            ocr_model = B200OcrModel(det_db_box_thresh=0.5, det_db_unclip_ratio=1.6, enable_merge_det_boxes=False, device=device, precision=precision)
        self.model = B200RapidTableModel(ocr_engine=ocr_model, device=device)

    def batch_predict(self, image_list, fill_image_res_list=None, **kwargs):
        fills = fill_image_res_list or [None] * len(image_list)
        return [self.model.predict(img, fill_image_res=fill) for img, fill in zip(image_list, fills)]


def make_ocr_model_init(device=0, precision=None, fallback=None, devices=None):
    """Factory with the signature of model_init.ocr_model_init (model_init.py:45-54).  devices=[0, 1, ...]: the model is a
    `B200OcrPool` that shards pages / text-line batches over those GPUs (rapiddoc_b200/multi.py)."""
    def ocr_model_init(det_db_box_thresh=0.5, lang=None, ocr_config=None, det_db_unclip_ratio=1.8, enable_merge_det_boxes=True,
                       is_seal=False):
        if is_seal:
            if fallback is None:
                raise NotImplementedError("seal OCR is outside the B200 hot path")
            return fallback(det_db_box_thresh, lang, ocr_config, det_db_unclip_ratio, enable_merge_det_boxes, is_seal)
        kw = dict(det_db_box_thresh=det_db_box_thresh, lang=lang, ocr_config=ocr_config, use_dilation=True,
                  det_db_unclip_ratio=det_db_unclip_ratio, enable_merge_det_boxes=enable_merge_det_boxes, precision=precision)
        if devices is not None and len(devices) > 1:
            from .multi import B200OcrPool
            return B200OcrPool(devices, **kw)
        from .ocr import B200OcrModel
        return B200OcrModel(device=devices[0] if devices else device, **kw)
    return ocr_model_init


def install(device=0, precision=None, devices=None):
    """Seam 2: rebind RapidDoc's OCR model factory.  Returns the original factory.  devices=[...] shards over several GPUs."""
    from rapid_doc.backend.pipeline import model_init as mi
    orig = mi.ocr_model_init
    mi.ocr_model_init = make_ocr_model_init(device, precision, fallback=orig, devices=devices)
    mi.AtomModelSingleton._models.clear()
    return orig


def install_window():
    """Seam 2b: rebind the two OCR callers of the batch scheduler (backend/pipeline/analyze_utils.py:105-292, imported by name
    into backend/pipeline/batch_analyze.py) to the window-level restatements of rapiddoc_b200/window.py: same data contract,
    but each size bucket is detected as one window.  Returns the original pair."""
    from rapid_doc.backend.pipeline import analyze_utils as au
    from rapid_doc.backend.pipeline.model_list import AtomicModel
    from . import window
    orig = (au._run_ocr_det_batch, au._run_ocr_rec_postprocess)

    def det(ocr_res_all_page, atom_model_manager, ocr_config):
        return window.run_ocr_det_batch(ocr_res_all_page, lambda lang: atom_model_manager.get_atom_model(
            atom_model_name=AtomicModel.OCR, lang=lang, ocr_config=ocr_config), ocr_config)

    def rec(images_layout_res, ocr_config):
        mgr = au.AtomModelSingleton()
        return window.run_ocr_rec_postprocess(images_layout_res, lambda lang: mgr.get_atom_model(
            atom_model_name=AtomicModel.OCR, lang=lang, ocr_config=ocr_config), ocr_config)
    au._run_ocr_det_batch, au._run_ocr_rec_postprocess = det, rec
    try:
        from rapid_doc.backend.pipeline import batch_analyze as ba
        ba._run_ocr_det_batch, ba._run_ocr_rec_postprocess = det, rec
    except Exception:
        pass
    return orig


def install_orientation(device=0):
    """Seam 2 for SURVEY T8: rebind the table-orientation classifier factory (`img_orientation_cls_model_init`,
    rapid_doc/backend/pipeline/model_init.py:32-34, -> RapidOrientationModel) to the CUDA classifier.  Returns the original."""
    from rapid_doc.backend.pipeline import model_init as mi
    from .orientation import B200OrientationModel
    orig = mi.img_orientation_cls_model_init
    mi.img_orientation_cls_model_init = lambda: B200OrientationModel(device=device)
    mi.AtomModelSingleton._models.clear()
    return orig


def install_table_structure(device=0):
    """Seam 3 for SURVEY T4: `RapidTable._init_table_structer` (rapid_table_self/main.py:64-76) returns the CUDA structurer for
    ModelType.SLANET1M — the wireless-table model RapidDoc ships — and the reference's own object for every other model type.
    Returns the original method."""
    from rapid_doc.model.table.rapid_table_self import main as rt
    from .table import B200TableStructurer
    orig = rt.RapidTable._init_table_structer

    def init_structurer(self):
        mt = getattr(self.cfg.model_type, "value", self.cfg.model_type)
        if mt == "slanet_1m":
            return B200TableStructurer(model_path=self.cfg.model_dir_or_path, model_type=mt, device=device)
        return orig(self)
    rt.RapidTable._init_table_structer = init_structurer
    return orig


def session_for(cfg):
    """Seam 3: pick the B200 session from the configured model file (det vs rec)."""
    from .engine import B200DetSession, B200RecSession
    path = str(cfg.get("model_path", "") if hasattr(cfg, "get") else "")
    return B200RecSession(cfg) if "rec" in path.lower() else B200DetSession(cfg)


def install_engine():
    """Seam 3: replace rapidocr's TorchInferSession (what ocr_patch.patch_torch_ocr does,
    rapid_doc/model/ocr/ocr_patch.py:95-105) with the B200 sessions."""
    try:
        from rapidocr.inference_engine.pytorch import main as rt
        import rapidocr.inference_engine.pytorch as rt_pkg
    except Exception:
        from rapidocr.inference_engine import torch as rt
        rt_pkg = None

    class _B200Session:
        def __new__(cls, cfg):
            return session_for(cfg)
    rt.TorchInferSession = _B200Session
    if rt_pkg is not None:
        rt_pkg.TorchInferSession = _B200Session
