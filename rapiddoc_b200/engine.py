"""Engine layer: thin Python handles over the C-ABI plus the two `InferSession`-protocol
classes that slot in where RapidDoc already swaps engines.

Reference seam: rapidocr's `InferSession` as replaced by rapid_doc/model/ocr/ocr_patch.py:95-105
with rapid_doc/model/ocr/torch.py:33-200 (`TorchInferSession.__call__(img) -> np.ndarray`,
`have_key`, `get_character_list`).  `B200DetSession` / `B200RecSession` keep exactly that
surface; the faster facade-level entry points (uint8 in, decoded ids out) are the
`DetEngine.infer_u8` / `RecEngine.infer_u8` methods used by rapiddoc_b200.ocr.
"""
import ctypes as C

import numpy as np

from . import _lib, weights as W

DET_MEAN = (0.485, 0.456, 0.406)   # rapid_doc/model/ocr/rapid_ocr.py:61-62
DET_STD = (0.229, 0.224, 0.225)


_NP2NAME = {np.dtype(np.uint8): "uint8", np.dtype(np.float32): "float32", np.dtype(np.int32): "int32"}


def _check_buf(name, a, dtype, numel, device):
    """A caller-supplied buffer goes to the C-ABI as a raw pointer: refuse anything whose dtype, contiguity, size or
    GPU does not match (a permuted / float page tensor or a tensor on another GPU would otherwise be silent garbage or an
    illegal address).  numpy arrays are host buffers; torch tensors may be host (pinned) or on the engine's device."""
    if a is None:
        return
    want = _NP2NAME[np.dtype(dtype)]
    if isinstance(a, np.ndarray):
        if a.dtype != np.dtype(dtype) or not a.flags.c_contiguous or a.size < numel:
            raise _lib.B200Error(f"{name}: need a C-contiguous {want} array of >= {numel} elements, got {a.dtype} {a.shape} "
                                 f"contiguous={a.flags.c_contiguous}")
        return
    if hasattr(a, "data_ptr"):
        got = str(a.dtype).replace("torch.", "")
        if got != want or not a.is_contiguous() or a.numel() < numel:
            raise _lib.B200Error(f"{name}: need a contiguous {want} tensor of >= {numel} elements, got {got} {tuple(a.shape)} "
                                 f"contiguous={a.is_contiguous()}")
        if a.is_cuda and a.device.index != device:
            raise _lib.B200Error(f"{name}: tensor lives on cuda:{a.device.index}, the engine on cuda:{device}")
        return
    raise _lib.B200Error(f"{name}: unsupported buffer type {type(a)}")


def _stream_ptr(stream):
    if stream is None:
        return None
    if isinstance(stream, int):
        return stream or None
    return getattr(stream, "cuda_stream", None) or None  # torch.cuda.Stream


class DetEngine:
    """PP-OCRv6-small DBNet on one GPU (rdb_det_*)."""

    def __init__(self, device=0, precision=_lib.PREC_FP16, blob=None, weights_path=None):
        self._lib = _lib.load()
        blob = blob if blob is not None else W.det_blob(weights_path)
        self._blob = blob
        h = C.c_void_p()
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        _lib.check(self._lib.rdb_det_create(buf, len(blob), int(device), int(precision), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.precision = int(precision)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rdb_det_destroy(self._h)
            self._h = None

    __del__ = close

    def set_chunk_pixels(self, px):
        _lib.check(self._lib.rdb_det_set_chunk_pixels(self._h, int(px)))

    @property
    def last_launches(self):
        return int(self._lib.rdb_det_last_launches(self._h))

    def infer_f32(self, x, prob=None, stream=None):
        """x [n,3,h,w] f32 (numpy or device tensor) -> prob [n,1,h,w] f32 (same kind)."""
        n, c, h, w = x.shape
        assert c == 3
        if prob is None:
            if isinstance(x, np.ndarray):
                prob = np.empty((n, 1, h, w), np.float32)
            else:
                import torch
                prob = torch.empty((n, 1, h, w), dtype=torch.float32, device=x.device)
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.float32)
        _check_buf("x", x, np.float32, n * 3 * h * w, self.device)
        _check_buf("prob", prob, np.float32, n * h * w, self.device)
        _lib.check(self._lib.rdb_det_infer_f32(self._h, _lib.ptr(x), n, h, w, _lib.ptr(prob), _stream_ptr(stream)))
        return prob

    def infer_u8(self, pages, thresh=0.3, use_dilation=True, prob=None, bitmap=None, want_prob=True, want_bitmap=True,
                 mean=DET_MEAN, std=DET_STD, stream=None, resize_to=None):
        """pages [n,h,w,3] uint8 BGR (numpy / device tensor) -> (prob [n,h,w] f32, bitmap [n,h,w] u8).
        resize_to=(rh, rw): DetPreProcess' cv2.resize runs on the GPU first (bit-exact); outputs are [n,rh,rw]."""
        n, h, w, c = pages.shape
        assert c == 3
        src_h, src_w = h, w
        if resize_to is not None:
            h, w = int(resize_to[0]), int(resize_to[1])
        is_np = isinstance(pages, np.ndarray)
        if is_np:
            pages = np.ascontiguousarray(pages, dtype=np.uint8)
        if prob is None and want_prob:
            if is_np:
                prob = np.empty((n, h, w), np.float32)
            else:
                import torch
                prob = torch.empty((n, h, w), dtype=torch.float32, device=pages.device)
        if bitmap is None and want_bitmap:
            if is_np:
                bitmap = np.empty((n, h, w), np.uint8)
            else:
                import torch
                bitmap = torch.empty((n, h, w), dtype=torch.uint8, device=pages.device)
        _check_buf("pages", pages, np.uint8, n * src_h * src_w * 3, self.device)
        _check_buf("prob", prob, np.float32, n * h * w, self.device)
        _check_buf("bitmap", bitmap, np.uint8, n * h * w, self.device)
        m = (C.c_float * 3)(*mean)
        s = (C.c_float * 3)(*std)
        if resize_to is not None and (src_h, src_w) != (h, w):
            _lib.check(self._lib.rdb_det_infer_u8_resize(self._h, _lib.ptr(pages), n, src_h, src_w, h, w, m, s, float(thresh),
                                                         int(bool(use_dilation)), _lib.ptr(prob), _lib.ptr(bitmap), _stream_ptr(stream)))
        else:
            _lib.check(self._lib.rdb_det_infer_u8(self._h, _lib.ptr(pages), n, h, w, m, s, float(thresh), int(bool(use_dilation)),
                                                  _lib.ptr(prob), _lib.ptr(bitmap), _stream_ptr(stream)))
        return prob, bitmap


class RecEngine:
    """PP-OCRv6-small LightSVTR/CTC recogniser on one GPU (rdb_rec_*)."""

    def __init__(self, device=0, precision=_lib.PREC_FP16, blob=None, weights_path=None):
        self._lib = _lib.load()
        blob = blob if blob is not None else W.rec_blob(weights_path)
        self._blob = blob
        h = C.c_void_p()
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        _lib.check(self._lib.rdb_rec_create(buf, len(blob), int(device), int(precision), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.precision = int(precision)
        self.vocab = int(self._lib.rdb_rec_vocab(h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rdb_rec_destroy(self._h)
            self._h = None

    __del__ = close

    def set_chunk_crops(self, n):
        _lib.check(self._lib.rdb_rec_set_chunk_crops(self._h, int(n)))

    @property
    def last_launches(self):
        return int(self._lib.rdb_rec_last_launches(self._h))

    def tokens(self, width):
        return int(self._lib.rdb_rec_tokens(int(width)))

    def _outs(self, n, T, like, want_softmax):
        if isinstance(like, np.ndarray):
            mk = lambda shape, dt: np.empty(shape, dt)
            i32, f32 = np.int32, np.float32
        else:
            import torch
            mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=like.device)
            i32, f32 = torch.int32, torch.float32
        o = dict(ids=mk((n, T), i32), probs=mk((n, T), f32), text_ids=mk((n, T), i32), text_len=mk((n,), i32), conf=mk((n,), f32))
        o["softmax"] = mk((n, T, self.vocab), f32) if want_softmax else None
        return o

    def infer_f32(self, x, want_softmax=False, stream=None):
        """x [n,3,48,w] f32 -> dict(ids, probs, text_ids, text_len, conf[, softmax])."""
        n, c, h, w = x.shape
        assert c == 3 and h == 48
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.float32)
        o = self._outs(n, self.tokens(w), x, want_softmax)
        _check_buf("x", x, np.float32, n * 3 * 48 * w, self.device)
        _lib.check(self._lib.rdb_rec_infer_f32(self._h, _lib.ptr(x), n, w, _lib.ptr(o["ids"]), _lib.ptr(o["probs"]),
                                               _lib.ptr(o["text_ids"]), _lib.ptr(o["text_len"]), _lib.ptr(o["conf"]),
                                               _lib.ptr(o["softmax"]), _stream_ptr(stream)))
        return o

    def infer_u8(self, crops, valid_w=None, stream=None, outs=None):
        """crops [n,48,w,3] uint8 BGR, valid_w [n] int32 -> dict(ids, probs, text_ids, text_len, conf)."""
        n, h, w, c = crops.shape
        assert c == 3 and h == 48
        if isinstance(crops, np.ndarray):
            crops = np.ascontiguousarray(crops, dtype=np.uint8)
            if valid_w is not None:
                valid_w = np.ascontiguousarray(valid_w, dtype=np.int32)
        T = self.tokens(w)
        o = outs if outs is not None else self._outs(n, T, crops, False)
        _check_buf("crops", crops, np.uint8, n * 48 * w * 3, self.device)
        _check_buf("valid_w", valid_w, np.int32, n, self.device)
        for k, dt, ne in (("ids", np.int32, n * T), ("probs", np.float32, n * T), ("text_ids", np.int32, n * T), ("text_len", np.int32, n),
                          ("conf", np.float32, n)):
            _check_buf(k, o.get(k), dt, ne, self.device)
        _lib.check(self._lib.rdb_rec_infer_u8(self._h, _lib.ptr(crops), _lib.ptr(valid_w), n, w, _lib.ptr(o["ids"]),
                                              _lib.ptr(o["probs"]), _lib.ptr(o["text_ids"]), _lib.ptr(o["text_len"]),
                                              _lib.ptr(o["conf"]), _stream_ptr(stream)))
        return o

    def infer_u8_raw(self, crops_ptr, valid_w_ptr, n, w, ids_ptr, probs_ptr, stream=None):
        """rdb_rec_infer_u8 on raw DEVICE addresses (ints): crops [n,48,w,3] u8, valid_w [n] i32 -> ids / probs [n,T].
        Asynchronous on `stream`; used by the window-level recogniser, which carves all of these out of a few big buffers."""
        _lib.check(self._lib.rdb_rec_infer_u8(self._h, int(crops_ptr), int(valid_w_ptr), int(n), int(w), int(ids_ptr), int(probs_ptr),
                                              None, None, None, _stream_ptr(stream)))


# ------------------------------------------------------------------ InferSession protocol
class _Cfg(dict):
    __getattr__ = dict.get


class B200DetSession:
    """Drop-in for rapidocr's det `InferSession` (torch.py:33-200 shape): np [B,3,H,W] f32 ->
    np [B,1,H,W] f32 prob map."""

    def __init__(self, cfg=None):
        cfg = _Cfg(cfg or {})
        eng = cfg.get("engine_cfg") or {}
        self.engine = DetEngine(device=int(eng.get("gpu_id", 0) or 0), precision=int(cfg.get("precision", _lib.PREC_FP16)),
                                weights_path=cfg.get("model_path"))

    def __call__(self, img: np.ndarray) -> np.ndarray:
        return self.engine.infer_f32(np.asarray(img, dtype=np.float32))

    def have_key(self, key: str = "character") -> bool:   # torch.py:194-195
        return False

    def get_character_list(self, key: str = "character"):  # torch.py:197-198
        return []


class B200RecSession:
    """Drop-in for rapidocr's rec `InferSession`: np [B,3,48,W] f32 -> np [B,T,V] f32 softmax
    probabilities (torch.py:186-192).  `decode()` is the fused fast path that skips the
    [B,T,V] tensor."""

    def __init__(self, cfg=None):
        cfg = _Cfg(cfg or {})
        eng = cfg.get("engine_cfg") or {}
        self.engine = RecEngine(device=int(eng.get("gpu_id", 0) or 0), precision=int(cfg.get("precision", _lib.PREC_FP16)),
                                weights_path=cfg.get("model_path"))
        self._characters = None

    def __call__(self, img: np.ndarray) -> np.ndarray:
        return self.engine.infer_f32(np.asarray(img, dtype=np.float32), want_softmax=True)["softmax"]

    def decode(self, img: np.ndarray):
        return self.engine.infer_f32(np.asarray(img, dtype=np.float32))

    def have_key(self, key: str = "character") -> bool:
        return False

    def get_character_list(self, key: str = "character"):
        return []
