"""e2e det timing with host buffers: with / without the fp32 prob map coming back (is D2H the bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rapiddoc_b200 import PREC_FP16, synth
from rapiddoc_b200.engine import DetEngine
B = 64
base = synth.det_pages(8, 1024, 1024, seed=1)
host = torch.empty((B, 1024, 1024, 3), dtype=torch.uint8).pin_memory()
for i in range(B):
    host[i] = torch.from_numpy(base[i % 8])
hp = torch.empty((B, 1024, 1024), dtype=torch.float32).pin_memory().numpy()
hb = torch.empty((B, 1024, 1024), dtype=torch.uint8).pin_memory().numpy()
eng = DetEngine(0, PREC_FP16)
hn = host.numpy()
def run(fn, tag):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"{tag:34s} {dt*1e3:7.2f} ms/step {B/dt:8.1f} pages/s")
run(lambda: eng.infer_u8(hn, prob=hp, bitmap=hb), "prob f32 + bitmap (bench e2e)")
run(lambda: eng.infer_u8(hn, bitmap=hb, want_prob=False), "bitmap only")
for pages in (12, 16, 24, 32):
    for ramp in (0, 1, 2):
        os.environ["RDB_RAMP"] = str(ramp)
        eng.set_chunk_pixels(pages * 1024 * 1024)
        run(lambda: eng.infer_u8(hn, prob=hp, bitmap=hb), f"host chunk={pages} ramp={ramp}")
del os.environ["RDB_RAMP"]
eng.set_chunk_pixels(32 * 1024 * 1024)
d = host.cuda()
dp = torch.empty((B, 1024, 1024), dtype=torch.float32, device="cuda"); db = torch.empty((B, 1024, 1024), dtype=torch.uint8, device="cuda")
run(lambda: eng.infer_u8(d, prob=dp, bitmap=db), "device resident")
# raw copy rates
for name, src, dst in (("H2D pages 201MB", host, d), ("D2H prob 268MB", dp, torch.from_numpy(hp))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"{name:34s} {src.numel()*src.element_size()/dt/1e9:6.1f} GB/s")
