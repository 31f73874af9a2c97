"""Error statistics of the fp16 CUDA path against the CPU oracle for a list of environment variants.
usage: python tools/err_stats.py "" "RDB_DW7_ROWS=2" "RDB_DW3=h2" ...   (each argument: space-separated NAME=VALUE pairs)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import nets, ocr_post as P
from rapiddoc_b200 import PREC_FP16
from rapiddoc_b200.engine import DetEngine, RecEngine

rng = np.random.default_rng(5)
pages = rng.integers(0, 256, (2, 288, 352, 3), dtype=np.uint8)
pages[:, 60:140, 40:200] = 250
pages[:, 150:, :] = (pages[:, 150:, :] // 32) * 32
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "det_real_192x256.npz"))
real = g["page_bgr"][None]
want = np.stack([nets.det_forward(P.det_preprocess(p, limit_side_len=4096))[0, 0] for p in pages])
want_real = g["prob"][0, 0][None]
x = rng.standard_normal((8, 3, 48, 320)).astype(np.float32)
logits = nets.rec_logits(x)
srt = np.sort(logits, axis=2)
margin = srt[:, :, -1] - srt[:, :, -2]
det, rec = DetEngine(0, PREC_FP16), RecEngine(0, PREC_FP16)
for variant in sys.argv[1:] or [""]:
    kv = dict(p.split("=") for p in variant.split()) if variant else {}
    os.environ.update(kv)
    try:
        prob, _ = det.infer_u8(pages)
        prob_r, _ = det.infer_u8(real)
        out = rec.infer_f32(x)
    finally:
        for k in kv:
            del os.environ[k]
    d, dr = np.abs(prob - want), np.abs(prob_r - want_real)
    flips = ((prob > 0.3) != (want > 0.3)).sum() + ((prob_r > 0.3) != (want_real > 0.3)).sum()
    mism = out["ids"] != logits.argmax(2)
    print(f"{variant or 'default':28s} det synth max {d.max():.2e} mean {d.mean():.2e} | real max {dr.max():.2e} mean {dr.mean():.2e} | "
          f"thr flips {flips} | rec argmax mismatches {mism.sum()} (max margin there {margin[mism].max() if mism.any() else 0:.3f})")
