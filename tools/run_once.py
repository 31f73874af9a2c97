"""Run the det (or rec) hot path a few times on one chunk-sized batch; used under ncu.
usage: python tools/run_once.py det|rec [batch] [passes]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapiddoc_b200 import PREC_FP16, synth  # noqa: E402
from rapiddoc_b200.engine import DetEngine, RecEngine  # noqa: E402
import torch  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "det"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else (8 if wl == "det" else 512)
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
if wl == "det":
    pages = torch.from_numpy(synth.det_pages(min(batch, 4), 1024, 1024, seed=1)).repeat((batch + 3) // 4, 1, 1, 1)[:batch].cuda()
    eng = DetEngine(0, PREC_FP16)
    for _ in range(passes):
        eng.infer_u8(pages, stream=torch.cuda.current_stream())
else:
    crops = torch.from_numpy(synth.rec_crops(batch, 48, 320, seed=2)).cuda()
    vw = torch.full((batch,), 320, dtype=torch.int32, device="cuda")
    eng = RecEngine(0, PREC_FP16)
    for _ in range(passes):
        eng.infer_u8(crops, vw, stream=torch.cuda.current_stream())
torch.cuda.synchronize()
print("done", eng.last_launches)
