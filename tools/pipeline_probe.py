"""Facade-level det+rec pipeline (B200OcrModel.ocr: GPU det -> host contours/unclip (OpenCV, as the reference) -> GPU warp/resize ->
GPU rec + CTC) on real and synthetic pages: pages/s of ONE Python thread, with the share of host post-processing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2, numpy as np
from rapiddoc_b200 import PREC_FP16, synth
from rapiddoc_b200.ocr import B200OcrModel

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "page_img5_e2e.npz"))
real = cv2.imdecode(g["png"], cv2.IMREAD_COLOR)
pages = {"real 704x960 (14 lines)": real, "synthetic 1024x1024": synth.det_pages(1, 1024, 1024, seed=1)[0]}
model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, precision=PREC_FP16)
for name, img in pages.items():
    for gpu_crop in (True, False):
        model.gpu_crop = gpu_crop
        for _ in range(3):
            res = model.ocr(img, det=True, rec=True)[0]
        t0 = time.perf_counter()
        n = 20
        for _ in range(n):
            res = model.ocr(img, det=True, rec=True)[0]
        dt = (time.perf_counter() - t0) / n
        t1 = time.perf_counter()
        for _ in range(n):
            model.text_detector(img)
        ddet = (time.perf_counter() - t1) / n
        print(f"{name:28s} gpu_crop={gpu_crop!s:5s} {len(res or [])} lines: {dt*1e3:7.2f} ms/page = {1/dt:6.1f} pages/s (det incl. host DBPostProcess {ddet*1e3:6.2f} ms)")
