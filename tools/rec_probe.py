"""rec throughput vs chunk size (device-resident crops)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rapiddoc_b200 import PREC_FP16, synth
from rapiddoc_b200.engine import RecEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = torch.from_numpy(synth.rec_crops(B, 48, 320, seed=2)).cuda()
vw = torch.full((B,), 320, dtype=torch.int32, device="cuda")
eng = RecEngine(0, PREC_FP16)
outs = eng._outs(B, eng.tokens(320), x, False)
for chunk in (64, 128, 171, 256, 512):
    eng.set_chunk_crops(chunk)
    for _ in range(3): eng.infer_u8(x, vw, stream=torch.cuda.current_stream(), outs=outs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.infer_u8(x, vw, stream=torch.cuda.current_stream(), outs=outs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"chunk {chunk:4d}: {ms:.3f} ms -> {B / ms * 1e3:9.1f} crops/s")
