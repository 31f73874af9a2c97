"""CPU tests: the C-ABI library loads and exports every symbol include/rapiddoc_b200.h
declares; host-only entry points behave; device entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rapiddoc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "rapiddoc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rdb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported_and_typed():
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes signature"
    for n in _lib.SYMBOLS:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"


def test_version_and_tokens():
    lib = _lib.load()
    assert lib.rdb_version() == 100
    assert lib.rdb_rec_tokens(320) == 40
    assert lib.rdb_rec_tokens(1081) == 135
    assert lib.rdb_rec_tokens(173) == 22


def test_no_gpu_fails_loudly():
    lib = _lib.load()
    if lib.rdb_device_count() > 0:
        pytest.skip("GPU present")
    h = C.c_void_p()
    blob = b"RDW1" + b"\0" * 64
    rc = lib.rdb_det_create(blob, len(blob), 0, 0, C.byref(h))
    assert rc == _lib.RDB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.rdb_last_error()
    with pytest.raises(_lib.B200Error):
        from rapiddoc_b200.engine import DetEngine
        DetEngine(blob=blob)


def test_clipper_offset_matches_oracle():
    """host-only C++ Clipper restatement vs the oracle's independent Python restatement."""
    from oracle.ocr_post import clipper_offset_round
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for _ in range(200):
        cx, cy = rng.uniform(50, 900, 2)
        w, h = rng.uniform(6, 400), rng.uniform(4, 60)
        ang = rng.uniform(-0.6, 0.6)
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        box = (np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) / 2) @ R.T + [cx, cy]
        if rng.random() < 0.3:
            box = box[::-1].copy()
        box = box.astype(np.float32)
        area = 0.5 * abs(np.dot(box[:, 0], np.roll(box[:, 1], 1)) - np.dot(box[:, 1], np.roll(box[:, 0], 1)))
        per = np.linalg.norm(box - np.roll(box, 1, axis=0), axis=1).sum()
        d = float(area * rng.choice([1.6, 1.8, 0.5]) / per)
        want = clipper_offset_round(box, d)
        xy = (C.c_double * 8)(*box.astype(np.float64).reshape(-1))
        out = (C.c_int64 * 2048)()
        n = lib.rdb_clipper_offset(xy, 4, d, out, 1024)
        assert n == len(want)
        got = np.frombuffer(out, dtype=np.int64)[: 2 * n].reshape(-1, 2)
        assert np.array_equal(got, want)
