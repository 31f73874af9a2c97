"""Generates tests/golden/onnx_cases.npz: outputs of OpenCV's DNN importer (cv2.dnn.readNetFromONNX — an implementation
independent of both oracle/onnx_ref.py and the CUDA executor) on the two ONNX files of weights/, for seeded inputs that the tests
regenerate.  Run in the build container:  python oracle/make_golden_onnx.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rapiddoc_b200 import synth                      # noqa: E402
from oracle import onnx_ref                          # noqa: E402


def orientation_inputs():
    """Four 224x224 network inputs: a synthetic text page in its four rotations, through the reference preprocessing."""
    page = synth.det_pages(1, 640, 480, seed=11, lines=24)[0]
    rots = [page, cv2.rotate(page, cv2.ROTATE_90_COUNTERCLOCKWISE), cv2.rotate(page, cv2.ROTATE_180), cv2.rotate(page, cv2.ROTATE_90_CLOCKWISE)]
    return rots, np.stack([onnx_ref.orientation_preprocess(r) for r in rots])


def seal_input(seed=0, size=320):
    img = synth.seal_image(seed, size)
    mean = np.array([0.485, 0.456, 0.406], np.float32)
    std = np.array([0.229, 0.224, 0.225], np.float32)
    return np.ascontiguousarray(((img.astype(np.float32) / 255.0 - mean) / std).transpose(2, 0, 1)[None], np.float32)


def table_inputs():
    """Three synthetic table crops (4x3 ruled, 6x4 ruled, 3x2 borderless) through TablePreprocess (T3)."""
    from rapiddoc_b200 import table
    imgs = [synth.table_image(0, 4, 3), synth.table_image(1, 6, 4, 360, 520), synth.table_image(2, 3, 2, 200, 300, lines=False)]
    x, shapes = table.TablePreprocess()(imgs)
    return imgs, np.asarray(x, np.float32), shapes


def main():
    out = {}
    net = cv2.dnn.readNetFromONNX(os.path.join(ROOT, "weights", "rapid_orientation.onnx"))
    _, xs = orientation_inputs()
    ys = []
    for x in xs:                                     # the importer bakes batch 1 into the flatten
        net.setInput(x[None])
        ys.append(net.forward().copy())
    out["orientation_scores"] = np.concatenate(ys)
    net = cv2.dnn.readNetFromONNX(os.path.join(ROOT, "weights", "pp-ocrv4_mobile_seal_det.onnx"))
    net.setInput(seal_input())
    out["seal_prob"] = net.forward().copy()[0, 0].astype(np.float16)      # fp16 storage: the tests compare at 2e-3
    # SLANet: cv2.dnn cannot import the graph (dynamic shapes + Loop), so these rows come from oracle/onnx_ref.py itself — a
    # regression fixture for the literal node-by-node execution of the file, not an independent pin
    _, x, _ = table_inputs()
    loc, probs = onnx_ref.run(os.path.join(ROOT, "weights", "slanet-1m.onnx"), x)
    out["slanet_ids"] = probs.argmax(-1).astype(np.int16)
    out["slanet_loc"] = loc.astype(np.float16)
    out["slanet_pmax"] = probs.max(-1).astype(np.float16)
    # the reference's DEFAULT engine runs these two ONNX files through onnxruntime; the repo's det / rec oracles (oracle/nets.py)
    # are torch restatements on the safetensors weights.  cv2.dnn on the det file and the node-by-node interpreter on the rec
    # file (cv2.dnn cannot import its dynamic shapes) give the model files' own outputs for seeded inputs
    ref_res = "/root/reference/rapid_doc/resources"
    if os.path.isdir(ref_res):
        from oracle import ocr_post as P
        page = synth.det_pages(1, 256, 512, seed=4, lines=5)[0]
        x = P.det_preprocess(page, limit_side_len=4096)
        net = cv2.dnn.readNetFromONNX(os.path.join(ref_res, "ch_PP-OCRv6_det_small.onnx"))
        net.setInput(x)
        out["ocr_det_onnx_prob"] = net.forward().copy()[0, 0].astype(np.float16)
        xr = np.random.RandomState(0).randn(3, 3, 48, 160).astype(np.float32)
        pr = onnx_ref.run(os.path.join(ref_res, "ch_PP-OCRv6_rec_small.onnx"), xr)
        out["ocr_rec_onnx_ids"] = pr.argmax(-1).astype(np.int32)
        out["ocr_rec_onnx_pmax"] = pr.max(-1).astype(np.float32)
    else:
        old = np.load(os.path.join(ROOT, "tests", "golden", "onnx_cases.npz"))
        for k in ("ocr_det_onnx_prob", "ocr_rec_onnx_ids", "ocr_rec_onnx_pmax"):
            out[k] = old[k]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "onnx_cases.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()}, out["orientation_scores"].round(3))


if __name__ == "__main__":
    main()
