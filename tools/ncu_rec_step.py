"""One warm recognition batch (256 crops 48x688, the pipeline workload's mean width) for ncu (`--profile-from-start off`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rapiddoc_b200 import PREC_FP16, synth  # noqa: E402
from rapiddoc_b200.engine import RecEngine  # noqa: E402

n, w = 256, 688
x = torch.from_numpy(synth.rec_crops(n, 48, w, seed=2)).cuda()
vw = torch.full((n,), w, dtype=torch.int32, device="cuda")
os.environ["RDB_LANES"] = "1"
eng = RecEngine(0, PREC_FP16)
for _ in range(2):
    eng.infer_u8(x, vw, stream=torch.cuda.current_stream())
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.infer_u8(x, vw, stream=torch.cuda.current_stream())
torch.cuda.synchronize()
torch.cuda.profiler.stop()
