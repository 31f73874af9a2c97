"""Per-kernel device times of one det / rec pass (rdb_profile_*).  usage: quick_prof.py det|rec [batch] [top]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from rapiddoc_b200 import PREC_FP16, _lib, synth  # noqa: E402
from rapiddoc_b200.engine import DetEngine, RecEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "det"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else (8 if wl == "det" else 512)
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
if wl == "det":
    x = torch.from_numpy(synth.det_pages(min(batch, 4), 1024, 1024, seed=1)).repeat((batch + 3) // 4, 1, 1, 1)[:batch].cuda()
    eng = DetEngine(0, PREC_FP16)
    run = lambda: eng.infer_u8(x, stream=torch.cuda.current_stream())
else:
    x = torch.from_numpy(synth.rec_crops(batch, 48, 320, seed=2)).cuda()
    vw = torch.full((batch,), 320, dtype=torch.int32, device="cuda")
    eng = RecEngine(0, PREC_FP16)
    run = lambda: eng.infer_u8(x, vw, stream=torch.cuda.current_stream())
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{wl} batch {batch}: {ms:.3f} ms/pass -> {batch / ms * 1e3:.1f} units/s")
_lib.profile(True); _lib.profile_reset()
for _ in range(3):
    run()
torch.cuda.synchronize()
prof = _lib.profile_dump(); _lib.profile(False)
tot = sum(v[0] for v in prof.values())
print(f"sum of kernel times {tot / 3:.3f} ms/pass")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {k:52s} n={v[1]//3:3d} avg_us={v[0]/v[1]*1e3:8.1f} share={100*v[0]/tot:5.1f}%")
