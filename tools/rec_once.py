import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rapiddoc_b200 import PREC_FP16, synth
from rapiddoc_b200.engine import RecEngine
x = torch.from_numpy(synth.rec_crops(256, 48, 320, seed=2)).cuda()
vw = torch.full((256,), 320, dtype=torch.int32, device="cuda")
eng = RecEngine(0, PREC_FP16)
eng.infer_u8(x, vw, stream=torch.cuda.current_stream()); torch.cuda.synchronize()
