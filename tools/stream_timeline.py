"""Wall-clock timeline of the two stages of ocr_pages_stream (diagnostic)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rapiddoc_b200  # sets CUDA_DEVICE_MAX_CONNECTIONS before torch creates the context
import torch
from rapiddoc_b200 import synth, dbpost
from rapiddoc_b200.ocr import B200OcrModel
n = 64
base = synth.det_pages(8, 1024, 1024, seed=1)
pages = torch.stack([torch.from_numpy(np.roll(base[i % 8], shift=(7 * (i // 8), 13 * (i // 8)), axis=(0, 1))) for i in range(n)]).cuda()
model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, ocr_config={"Det.limit_side_len": 1024, "Rec.rec_batch_num": 256})
for _ in range(3):
    model.ocr_pages(pages)
T0 = time.perf_counter()
log = []
wc, wt = model._window_crops, model._window_texts
def a(*args, **kw):
    t = time.perf_counter(); r = wc(*args, **kw); log.append(("A", t - T0, time.perf_counter() - T0)); return r
def b(*args, **kw):
    t = time.perf_counter(); r = wt(*args, **kw); log.append(("B", t - T0, time.perf_counter() - T0)); return r
model._window_crops, model._window_texts = a, b
dbpost.TIMES.clear()
t = time.perf_counter()
for _ in model.ocr_pages_stream(pages for _ in range(6)):
    pass
torch.cuda.synchronize()
print("total per step ms", (time.perf_counter() - t) / 6 * 1e3)
for k, s, e in sorted(log, key=lambda x: x[1]):
    print(f"{k} {s*1e3:8.1f} -> {e*1e3:8.1f}  ({(e-s)*1e3:6.1f} ms)")
for k, v in sorted(dbpost.TIMES.items()):
    print(f"  {k:40s} {v / 6 * 1e3:8.1f} ms")
