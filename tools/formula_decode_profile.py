"""Per-kernel device time of the formula decoder (non-graph mode, events around every op)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rapiddoc_b200
import torch
from rapiddoc_b200 import _lib, formula as FM, PREC_FP16
sd = FM.synthetic_state_dict(seed=0)
eng = FM.FormulaEngine(sd, precision=PREC_FP16, max_new_tokens=16, sync_every=64, use_graph=False)
x = torch.randn(32, 1, 384, 384, device="cuda")
enc = eng.encode(x)
eng.generate(enc)
torch.cuda.synchronize()
_lib.profile(True); _lib.profile_reset()
eng.generate(enc)
torch.cuda.synchronize()
prof = _lib.profile_dump(); _lib.profile(False)
tot = sum(v[0] for v in prof.values())
print("decoder kernel time per step (16 steps): %.3f ms" % (tot / 16))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:16]:
    print(f"{k:44s} {v[0]/16*1e3:8.1f} us/step {v[1]/16:6.1f} launches/step  avg {v[0]/v[1]*1e3:6.1f} us")
