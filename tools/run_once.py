"""one det pass of n pages (debug helper): python tools/run_once.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rapiddoc_b200 import PREC_FP16, synth
from rapiddoc_b200.engine import DetEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = torch.from_numpy(synth.det_pages(min(n, 4), 1024, 1024, seed=1)).repeat((n + 3) // 4, 1, 1, 1)[:n].cuda()
eng = DetEngine(0, PREC_FP16)
eng.infer_u8(x, stream=torch.cuda.current_stream())
torch.cuda.synchronize()
