"""Where the wall clock of one `B200OcrModel.ocr_pages` window goes (host side): cProfile of a warm step.
python tools/pipeline_profile.py [pages] [rec_batch]"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rapiddoc_b200 import synth  # noqa: E402
from rapiddoc_b200.ocr import B200OcrModel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
base = synth.det_pages(8, 1024, 1024, seed=1)
host = torch.empty((n, 1024, 1024, 3), dtype=torch.uint8).pin_memory()
for i in range(n):
    host[i] = torch.from_numpy(np.roll(base[i % 8], shift=(7 * (i // 8), 13 * (i // 8)), axis=(0, 1)))
pages = [host.numpy()[i] for i in range(n)]
model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, ocr_config={"Det.limit_side_len": 1024, "Rec.rec_batch_num": rb})
for _ in range(3):
    model.ocr_pages(pages)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    model.ocr_pages(pages)
torch.cuda.synchronize()
print(f"ocr_pages: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per {n} pages")
from rapiddoc_b200 import dbpost  # noqa: E402
dbpost.TIMES.clear()
for _ in range(3):
    model.ocr_pages(pages)
torch.cuda.synchronize()
for k, v in sorted(dbpost.TIMES.items()):
    print(f"  {k:40s} {v / 3 * 1e3:8.1f} ms")
t0 = time.perf_counter()
for _ in range(3):
    model.det_batch_predict(pages)
print(f"det_batch_predict: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per {n} pages")
pr = cProfile.Profile()
pr.enable()
model.ocr_pages(pages)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue())
