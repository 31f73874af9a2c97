"""ctypes binding of librapiddoc_b200.so (the C-ABI in include/rapiddoc_b200.h).

There is NO CPU fallback: if the shared library is missing or no B200 is visible the
constructors raise.  The library is built in-tree by rapiddoc_b200/build.py.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librapiddoc_b200.so")

RDB_OK = 0
RDB_ERR_INVALID, RDB_ERR_CUDA, RDB_ERR_NO_DEVICE = -1, -2, -3
PREC_FP32, PREC_FP16, PREC_TF32 = 0, 1, 2

# every symbol include/rapiddoc_b200.h declares: name -> (restype, argtypes)
_vp, _i, _f = C.c_void_p, C.c_int, C.c_float
SYMBOLS = {
    "rdb_version": (_i, []),
    "rdb_last_error": (C.c_char_p, []),
    "rdb_device_count": (_i, []),
    "rdb_pinned_alloc": (_i, [C.c_size_t, C.POINTER(_vp)]),
    "rdb_pinned_free": (_i, [_vp]),
    "rdb_det_create": (_i, [_vp, C.c_size_t, _i, _i, C.POINTER(_vp)]),
    "rdb_det_destroy": (None, [_vp]),
    "rdb_det_infer_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "rdb_det_infer_u8": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _f, _i, _vp, _vp, _vp]),
    "rdb_det_infer_u8_resize": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _f, _i, _vp, _vp, _vp]),
    "rdb_resize_linear_u8": (_i, [_i, _vp, _i, _i, _i, _vp, _i, _i, _vp]),
    "rdb_db_bitmap": (_i, [_i, _vp, _i, _i, _i, _f, _i, _vp, _vp]),
    "rdb_warp_crops": (_i, [_i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, C.c_int64, _vp]),
    "rdb_resize_pack_u8": (_i, [_i, _vp, C.c_int64, _i, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "rdb_warp_crops_batch": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, _vp]),
    "rdb_resize_pack_slots": (_i, [_i, _vp, C.c_int64, _i, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, _i, _vp]),
    "rdb_db_box_scores": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "rdb_debug_fill_quad": (_i, [_vp, _i, _i, _vp]),
    "rdb_det_set_pool_cap_bytes": (_i, [_vp, C.c_size_t]),
    "rdb_rec_set_pool_cap_bytes": (_i, [_vp, C.c_size_t]),
    "rdb_det_pool_bytes": (C.c_longlong, [_vp]),
    "rdb_rec_pool_bytes": (C.c_longlong, [_vp]),
    "rdb_layout_nms": (_i, [_i, _vp, _i, _vp, _vp, _i, _f, _f, _vp, _vp, _vp]),
    "rdb_layout_containment": (_i, [_i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rdb_argmax_rows": (_i, [_i, _vp, C.c_longlong, _i, _vp, _vp, _vp]),
    "rdb_ops_last_error": (C.c_char_p, []),
    "rdb_op_gemm": (_i, [_i, _i, _vp, _i, C.c_longlong, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, C.c_longlong]),
    "rdb_op_conv_tc": (_i, [_i, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    "rdb_op_im2col": (_i, [_i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "rdb_op_dwconv": (_i, [_i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "rdb_op_maxpool2x2s1": (_i, [_i, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "rdb_op_copy_cols": (_i, [_i, _i, _i, _vp, C.c_longlong, _i, _i, _vp, _i, _i, _vp]),
    "rdb_op_layernorm": (_i, [_i, _vp, C.c_longlong, _i, _vp, _vp, _f, _vp, _vp]),
    "rdb_op_embed": (_i, [_i, _vp, _i, _i, _vp, _f, _vp, _i, _vp, _vp, _vp]),
    "rdb_op_attn_decode": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rdb_op_add": (_i, [_i, _vp, _vp, _vp, C.c_longlong, _vp]),
    "rdb_op_chain": (_i, [_i, _vp, C.c_longlong, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _vp]),
    "rdb_op_global_avgpool": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "rdb_op_mul_gate": (_i, [_i, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "rdb_op_resize_nearest": (_i, [_i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "rdb_op_depth_to_space": (_i, [_i, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "rdb_op_softmax_rows": (_i, [_i, _vp, C.c_longlong, _i, _vp, _vp]),
    "rdb_op_conv3x3_c4": (_i, [_i, _vp, _i, _i, _i, _vp, _i, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    "rdb_op_lut_u8_nhwc4": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "rdb_sla_last_error": (C.c_char_p, []),
    "rdb_sla_decode": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rdb_op_greedy_step": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "rdb_contours_trace": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_vp)]),
    "rdb_contours_counts": (_i, [_vp, _vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "rdb_contours_fetch": (_i, [_vp, _vp, _vp]),
    "rdb_contours_free": (None, [_vp]),
    "rdb_lines_sort_merge": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "rdb_debug_cubic_tab": (_i, [_vp]),
    "rdb_clipper_offset": (_i, [C.POINTER(C.c_double), _i, C.c_double, C.POINTER(C.c_int64), _i]),
    "rdb_clipper_offset_batch": (_i, [_vp, _i, _vp, _vp, _i, _vp]),
    "rdb_rec_create": (_i, [_vp, C.c_size_t, _i, _i, C.POINTER(_vp)]),
    "rdb_rec_destroy": (None, [_vp]),
    "rdb_rec_vocab": (_i, [_vp]),
    "rdb_rec_tokens": (_i, [_i]),
    "rdb_rec_infer_f32": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rdb_rec_infer_u8": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rdb_det_last_launches": (C.c_longlong, [_vp]),
    "rdb_rec_last_launches": (C.c_longlong, [_vp]),
    "rdb_debug_gemm": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "rdb_switches_reload": (_i, []),
    "rdb_profile_enable": (_i, [_i]),
    "rdb_profile_reset": (_i, []),
    "rdb_profile_dump": (_i, [C.c_char_p, C.c_size_t]),
    "rdb_det_set_chunk_pixels": (_i, [_vp, C.c_longlong]),
    "rdb_rec_set_chunk_crops": (_i, [_vp, _i]),
}

_lib = None


class B200Error(RuntimeError):
    pass


def load():
    """dlopen the in-tree library and type every exported symbol; raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(f"{LIB_PATH} not found — run `python -m rapiddoc_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check_op(rc):
    if rc < 0:
        raise B200Error(f"rapiddoc_b200 op error {rc}: {load().rdb_ops_last_error().decode('utf-8', 'replace')}")
    return rc


def check(rc):
    if rc < 0:
        msg = load().rdb_last_error().decode("utf-8", "replace")
        raise B200Error(f"rapiddoc_b200 error {rc}: {msg}")
    return rc


def ptr(a):
    """host numpy array / torch tensor (host or device) / int / None -> c_void_p value."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "ctypes"):          # numpy
        return a.ctypes.data
    if hasattr(a, "data_ptr"):        # torch
        return a.data_ptr()
    raise TypeError(type(a))


def profile(on=True):
    load().rdb_profile_enable(int(bool(on)))


def profile_reset():
    load().rdb_profile_reset()


def profile_dump():
    """{kernel name: (total_ms, launches)} accumulated since the last reset."""
    import json
    lib = load()
    n = lib.rdb_profile_dump(None, 0)
    buf = C.create_string_buffer(n)
    lib.rdb_profile_dump(buf, n)
    return {k: (v[0], int(v[1])) for k, v in json.loads(buf.value.decode()).items()}
