"""One warm SLANet step (32 synthetic tables, preprocessed, resident) for ncu: run under `ncu --profile-from-start off ...`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from rapiddoc_b200 import table  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ts = table.B200TableStructurer(device=0)
x, shapes = ts.preprocess_op(bench.table_inputs(n))
x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.float32))).cuda()
for _ in range(2):
    ts.session(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ts.session(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
