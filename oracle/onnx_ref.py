"""ORACLE (test infrastructure, never on the product path): a node-by-node CPU interpreter (torch fp32, NCHW, nothing fused or
folded) for the ONNX graphs RapidDoc runs through onnxruntime — `OrtInferSession.__call__`
(rapid_doc/model/orientation/rapid_orientation/utils.py:47-52) on rapid_orientation.onnx and the seal detector
pp-ocrv4_mobile_seal_det.onnx.  onnxruntime is not installed here, so each node follows the published ONNX operator definition
(opset 11-14: Conv, ConvTranspose, BatchNormalization inference form, HardSigmoid = max(0, min(1, alpha*x + beta)),
HardSwish = x * HardSigmoid(x; 1/6, 0.5), Resize nearest/asymmetric/floor, GlobalAveragePool, Softmax, ...).
Pinned: tests/test_onnx_run.py checks it against OpenCV's DNN importer (`cv2.dnn.readNetFromONNX`, an independent
implementation) run on the same files, and against tests/golden/onnx_cases.npz made from those cv2.dnn runs by
oracle/make_golden_onnx.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this file."""
import numpy as np
import torch
import torch.nn.functional as F

from rapiddoc_b200 import onnx_lite       # the protobuf reader only (no compute)


def run(path, x):
    """x: numpy [n,3,h,w] float32 -> numpy output of the graph."""
    g = onnx_lite.load(path)
    env = {k: torch.from_numpy(np.array(v)) for k, v in g.init.items()}
    env[g.inputs[0]] = torch.from_numpy(np.ascontiguousarray(x, np.float32))
    with torch.no_grad():
        for n in g.nodes:
            i = [env[k] if k else None for k in n.inputs]
            a = n.attrs
            if n.op == "Identity":
                y = i[0]
            elif n.op == "Conv":
                p = a.get("pads", [0, 0, 0, 0])
                assert p[0] == p[2] and p[1] == p[3]
                y = F.conv2d(i[0], i[1], i[2] if len(i) > 2 else None, a.get("strides", [1, 1]), (p[0], p[1]), a.get("dilations", [1, 1]), a.get("group", 1))
            elif n.op == "ConvTranspose":
                p = a.get("pads", [0, 0, 0, 0])
                y = F.conv_transpose2d(i[0], i[1], i[2] if len(i) > 2 else None, a.get("strides", [1, 1]), (p[0], p[1]), 0, a.get("group", 1))
            elif n.op == "BatchNormalization":
                y = F.batch_norm(i[0], i[3], i[4], i[1], i[2], False, 0.0, a.get("epsilon", 1e-5))
            elif n.op == "Relu":
                y = F.relu(i[0])
            elif n.op == "Sigmoid":
                y = torch.sigmoid(i[0])
            elif n.op == "HardSigmoid":
                y = torch.clamp(a.get("alpha", 0.2) * i[0] + a.get("beta", 0.5), 0.0, 1.0)
            elif n.op == "HardSwish":
                y = i[0] * torch.clamp(i[0] / 6.0 + 0.5, 0.0, 1.0)
            elif n.op == "Add":
                y = i[0] + i[1]
            elif n.op == "Mul":
                y = i[0] * i[1]
            elif n.op == "GlobalAveragePool":
                y = i[0].mean(dim=(2, 3), keepdim=True)
            elif n.op == "Resize":
                assert a.get("mode") == "nearest" and a.get("coordinate_transformation_mode") == "asymmetric" and a.get("nearest_mode", "floor") == "floor"
                s = i[2].numpy()
                oh, ow = int(i[0].shape[2] * s[2]), int(i[0].shape[3] * s[3])
                iy = torch.floor(torch.arange(oh) / float(s[2])).long().clamp(max=i[0].shape[2] - 1)
                ix = torch.floor(torch.arange(ow) / float(s[3])).long().clamp(max=i[0].shape[3] - 1)
                y = i[0][:, :, iy][:, :, :, ix]
            elif n.op == "Concat":
                y = torch.cat([t.reshape(-1) if t.dim() == 0 else t for t in i], dim=a.get("axis", 0))
            elif n.op == "Shape":
                y = torch.tensor(list(i[0].shape), dtype=torch.int64)
            elif n.op == "Slice":
                st, en = int(i[1].reshape(-1)[0]), int(i[2].reshape(-1)[0])
                ax = int(i[3].reshape(-1)[0]) if len(i) > 3 else 0
                y = i[0].narrow(ax, st, min(en, i[0].shape[ax]) - st)
            elif n.op == "Reshape":
                y = i[0].reshape([int(v) for v in i[1]])
            elif n.op == "MatMul":
                y = i[0] @ i[1]
            elif n.op == "Softmax":
                y = torch.softmax(i[0], dim=a.get("axis", -1))
            else:
                raise NotImplementedError(n.op)
            env[n.outputs[0]] = y
    return env[g.outputs[0]].numpy()


def orientation_preprocess(img):
    """RapidOrientation's PreProcess list (rapid_orientation/config.yaml: ResizeImage resize_short 256 -> CropImage 224 ->
    NormalizeImage -> ToCHWImage; utils.py:97-172), on the image exactly as RapidOrientationModel.predict passes it."""
    import cv2
    h, w = img.shape[:2]
    pct = float(256) / min(w, h)
    img = cv2.resize(img, (int(round(w * pct)), int(round(h * pct))))
    h, w = img.shape[:2]
    if h < 224 or w < 224:
        raise ValueError("CropImage: image smaller than the crop")
    ws, hs = (w - 224) // 2, (h - 224) // 2
    img = img[hs:hs + 224, ws:ws + 224, :]
    mean = np.array([0.485, 0.456, 0.406]).reshape(1, 1, 3).astype("float32")
    std = np.array([0.229, 0.224, 0.225]).reshape(1, 1, 3).astype("float32")
    x = (np.array(img).astype(np.float32) * np.float32(1.0 / 255.0) - mean) / std
    return x.astype(np.float32).transpose((2, 0, 1))


def orientation(path, img):
    """RapidOrientation.__call__ (rapid_orientation.py:43-56): label string of the arg-max class."""
    g = onnx_lite.load(path)
    labels = g.meta["character"].splitlines()
    out = run(path, orientation_preprocess(img)[None]).squeeze()
    return labels[int(np.argmax(out))], out
