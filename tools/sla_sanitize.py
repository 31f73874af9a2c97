"""Small SLANet decode run for compute-sanitizer (memcheck / racecheck): 3 tables, loop capped at 12 steps."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import make_golden_onnx as MG  # noqa: E402
from rapiddoc_b200.table import SlaNetSession  # noqa: E402

_, x, _ = MG.table_inputs()
s = SlaNetSession(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights", "slanet-1m.onnx"), 0)
s.max_steps, s.eos = 12, 3
loc, probs = s(x)
print("ids", probs.argmax(-1).tolist())
