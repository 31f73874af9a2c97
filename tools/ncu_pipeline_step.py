"""One warm `ocr_pages` step of the pipeline workload for ncu (profiling starts only after the warm-up:
run under `ncu --profile-from-start off ...`).  python tools/ncu_pipeline_step.py [pages]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rapiddoc_b200 import synth  # noqa: E402
from rapiddoc_b200.ocr import B200OcrModel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
base = synth.det_pages(8, 1024, 1024, seed=1)
pages = torch.stack([torch.from_numpy(np.roll(base[i % 8], shift=(7 * (i // 8), 13 * (i // 8)), axis=(0, 1))) for i in range(n)]).cuda()
os.environ["RDB_LANES"] = "1"      # set before the library reads its switches
model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, ocr_config={"Det.limit_side_len": 1024, "Rec.rec_batch_num": 256})
for _ in range(2):
    model.ocr_pages(pages)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.ocr_pages(pages)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
