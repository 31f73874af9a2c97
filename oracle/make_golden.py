"""Generate tests/golden/*.npz by running the REFERENCE's own torch networks (imported
from /root/reference, see oracle/ref_loader.py) on real demo images and seeded inputs.
Run in the build container only:  python -m oracle.make_golden
Inputs are stored next to the outputs so the fixtures are self-contained on the GPU box.
"""
import os
import sys

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ocr_post as P          # noqa: E402
from oracle import ref_loader as R        # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    det, rec = R.det_net(), R.rec_net()
    img = cv2.imread(f"{R.REF}/demo/images/img_5.png")            # 720x960 BGR slide
    # ---- det golden 1: real crop 192x256 (uint8 in, DetPreProcess, reference prob map)
    crop = np.ascontiguousarray(img[120:312, 40:296])
    x = P.det_preprocess(crop, limit_side_len=960)
    with torch.no_grad():
        prob = det(torch.from_numpy(x))["maps"].numpy()
    np.savez_compressed(os.path.join(OUT, "det_real_192x256.npz"), page_bgr=crop, x=x.astype(np.float32), prob=prob)
    # ---- det golden 2: seeded gaussian input, batch 2, 64x96
    xr = np.random.default_rng(7).standard_normal((2, 3, 64, 96)).astype(np.float32)
    with torch.no_grad():
        pr = det(torch.from_numpy(xr))["maps"].numpy()
    np.savez_compressed(os.path.join(OUT, "det_randn_2x64x96.npz"), x=xr, prob=pr)
    # ---- full page through the reference net -> DB post-process (oracle restatement) -> line crops
    xf = P.det_preprocess(img)
    with torch.no_grad():
        pf = det(torch.from_numpy(xf))["maps"].numpy()
    boxes, scores = P.db_postprocess(pf, img.shape[:2], box_thresh=0.3, unclip_ratio=1.8)
    boxes = np.array(P.sorted_boxes(boxes))
    crops = []
    for b in boxes[:8]:
        pts = np.array(b, np.float32)
        w = int(max(np.linalg.norm(pts[0] - pts[1]), np.linalg.norm(pts[2] - pts[3])))
        h = int(max(np.linalg.norm(pts[0] - pts[3]), np.linalg.norm(pts[1] - pts[2])))
        M = cv2.getPerspectiveTransform(pts, np.float32([[0, 0], [w, 0], [w, h], [0, h]]))
        crops.append(cv2.warpPerspective(img, M, (w, h), borderMode=cv2.BORDER_REPLICATE, flags=cv2.INTER_CUBIC))
    order = np.argsort([c.shape[1] / c.shape[0] for c in crops])
    crops = [crops[i] for i in order[:6]]
    xb, ratio = P.rec_batch_tensor(crops)
    with torch.no_grad():
        logits = rec(torch.from_numpy(xb))["ctc_logits"]
        probs = torch.softmax(logits, dim=2).numpy()
    ids = probs.argmax(2).astype(np.int32)
    mx = probs.max(2).astype(np.float32)
    top2 = np.sort(logits.numpy(), axis=2)[:, :, -2:]
    margin = (top2[:, :, 1] - top2[:, :, 0]).astype(np.float32)
    texts = P.ctc_decode(probs, R.characters())
    save = {f"crop{i}": c for i, c in enumerate(crops)}
    np.savez_compressed(os.path.join(OUT, "rec_real_6lines.npz"), x=xb, ids=ids, maxprob=mx, margin=margin,
                        texts=np.array([t for t, _ in texts]), conf=np.array([c for _, c in texts], np.float64), **save)
    # DB post golden on the same page (prob from the reference net; post by the restatement)
    np.savez_compressed(os.path.join(OUT, "det_page_img5.npz"), prob=pf.astype(np.float32), boxes=boxes.astype(np.float32),
                        scores=np.array(scores, np.float64), shape=np.array(img.shape[:2]))
    # ---- rec golden 2: seeded gaussian input, odd width
    xr = np.random.default_rng(11).standard_normal((3, 3, 48, 173)).astype(np.float32)
    with torch.no_grad():
        lg = rec(torch.from_numpy(xr))["ctc_logits"].numpy()
    np.savez_compressed(os.path.join(OUT, "rec_randn_3x48x173.npz"), x=xr, ids=lg.argmax(2).astype(np.int32),
                        logits_max=lg.max(2).astype(np.float32),
                        lse=torch.logsumexp(torch.from_numpy(lg), 2).numpy().astype(np.float32))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    print([t for t in texts])


if __name__ == "__main__":
    main()


def make_e2e_golden():
    """tests/golden/page_img5_e2e.npz: the demo page (PNG bytes), the reference-net boxes and the texts the
    reference flow produces on them (crop -> sort by ratio -> batches of 6 -> rec -> CTC decode)."""
    from oracle import nets
    img = cv2.imread(f"{R.REF}/demo/images/img_5.png")
    g = np.load(os.path.join(OUT, "det_page_img5.npz"))
    crops = []
    for b in g["boxes"]:
        pts = np.array(b, np.float32)
        w = int(max(np.linalg.norm(pts[0] - pts[1]), np.linalg.norm(pts[2] - pts[3])))
        h = int(max(np.linalg.norm(pts[0] - pts[3]), np.linalg.norm(pts[1] - pts[2])))
        M = cv2.getPerspectiveTransform(pts, np.float32([[0, 0], [w, 0], [w, h], [0, h]]))
        c = cv2.warpPerspective(img, M, (w, h), borderMode=cv2.BORDER_REPLICATE, flags=cv2.INTER_CUBIC)
        crops.append(np.rot90(c) if c.shape[0] * 1.0 / c.shape[1] >= 2 else c)
    order = np.argsort([c.shape[1] / c.shape[0] for c in crops])
    res = [None] * len(crops)
    for b0 in range(0, len(crops), 6):
        idx = order[b0:b0 + 6]
        xb, _ = P.rec_batch_tensor([crops[i] for i in idx])
        out = P.ctc_decode(nets.rec_forward(xb), nets.load_characters())
        for j, i in enumerate(idx):
            res[i] = out[j]
    ok, enc = cv2.imencode(".png", img)
    np.savez_compressed(os.path.join(OUT, "page_img5_e2e.npz"), png=enc, texts=np.array([t for t, _ in res]),
                        conf=np.array([c for _, c in res]), boxes=g["boxes"])


if __name__ == "__main__" and "--e2e" in sys.argv:
    make_e2e_golden()
