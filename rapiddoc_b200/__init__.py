"""rapiddoc_b200 — B200-native (sm_100a) OCR hot path behind RapidDoc's model plugin surface."""
import os as _os

# The engines use > 8 CUDA streams per GPU (compute lanes, copy streams, the detection-stage and scorer streams): with the
# default of 8 hardware work queues, streams alias onto the same queue and a tiny high-priority kernel can serialise behind
# a whole recognition window enqueued earlier on another stream.  Must be set before the CUDA context exists; a value the
# user exported wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib  # noqa: F401,E402
from ._lib import B200Error, PREC_FP16, PREC_FP32  # noqa: F401,E402

__version__ = "0.1.0"
