"""rapiddoc_b200 — B200-native (sm_100a) OCR hot path behind RapidDoc's model plugin surface."""
from . import _lib  # noqa: F401
from ._lib import B200Error, PREC_FP16, PREC_FP32  # noqa: F401

__version__ = "0.1.0"
