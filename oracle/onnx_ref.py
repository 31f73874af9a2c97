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


_TORCH_DT = {1: torch.float32, 6: torch.int32, 7: torch.int64, 9: torch.bool, 10: torch.float16, 11: torch.float64}


def _t(v):
    return torch.from_numpy(np.array(v)) if isinstance(v, np.ndarray) else v


def _ints(t):
    return [int(v) for v in t.reshape(-1)]


def _exec(g, env):
    """Execute graph g node by node; `env` already holds its inputs and every outer-scope value (ONNX subgraph scoping)."""
    for k, v in g.init.items():
        env[k] = _t(v)
    for n in g.nodes:
        i = [env[k] if k else None for k in n.inputs]
        a = n.attrs
        ys = None
        if n.op == "Identity":
            y = i[0]
        elif n.op == "Constant":
            y = _t(np.asarray(a["value"]))
        elif n.op == "Conv":
            x, st = i[0], a.get("strides", [1, 1])
            if a.get("auto_pad", "NOTSET") == "SAME_UPPER":           # output = ceil(in / stride); the odd padding element goes at the end
                k = i[1].shape[2:]
                ph = max((-(-x.shape[2] // st[0]) - 1) * st[0] + k[0] - x.shape[2], 0)
                pw = max((-(-x.shape[3] // st[1]) - 1) * st[1] + k[1] - x.shape[3], 0)
                x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
                p = [0, 0, 0, 0]
            else:
                p = a.get("pads", [0, 0, 0, 0])
                assert p[0] == p[2] and p[1] == p[3]
            y = F.conv2d(x, i[1], i[2] if len(i) > 2 else None, st, (p[0], p[1]), a.get("dilations", [1, 1]), a.get("group", 1))
        elif n.op == "ConvTranspose":
            p = a.get("pads", [0, 0, 0, 0])
            y = F.conv_transpose2d(i[0], i[1], i[2] if len(i) > 2 else None, a.get("strides", [1, 1]), (p[0], p[1]), 0, a.get("group", 1))
        elif n.op == "BatchNormalization":
            y = F.batch_norm(i[0], i[3], i[4], i[1], i[2], False, 0.0, a.get("epsilon", 1e-5))
        elif n.op == "Relu":
            y = F.relu(i[0])
        elif n.op == "Sigmoid":
            y = torch.sigmoid(i[0])
        elif n.op == "Tanh":
            y = torch.tanh(i[0])
        elif n.op == "HardSigmoid":
            y = torch.clamp(a.get("alpha", 0.2) * i[0] + a.get("beta", 0.5), 0.0, 1.0)
        elif n.op == "HardSwish":
            y = i[0] * torch.clamp(i[0] / 6.0 + 0.5, 0.0, 1.0)
        elif n.op == "Add":
            y = i[0] + i[1]
        elif n.op == "Sub":
            y = i[0] - i[1]
        elif n.op == "Mul":
            y = i[0] * i[1]
        elif n.op == "Div":
            y = i[0] / i[1]
        elif n.op == "Pow":
            y = torch.pow(i[0], i[1])
        elif n.op == "Sqrt":
            y = torch.sqrt(i[0])
        elif n.op == "Erf":
            y = torch.erf(i[0])
        elif n.op == "ReduceMean":
            axes = [int(v) - (1 << 64) if int(v) >= (1 << 63) else int(v) for v in a["axes"]]
            y = i[0].mean(dim=axes, keepdim=bool(a.get("keepdims", 1)))
        elif n.op in ("MaxPool", "AveragePool"):
            k, st = a["kernel_shape"], a.get("strides", a["kernel_shape"])
            x = i[0]
            if a.get("auto_pad", "NOTSET") == "SAME_UPPER":
                ph = max((-(-x.shape[2] // st[0]) - 1) * st[0] + k[0] - x.shape[2], 0)
                pw = max((-(-x.shape[3] // st[1]) - 1) * st[1] + k[1] - x.shape[3], 0)
                pads = (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)                  # extra padding at the end (UPPER)
            else:
                p4 = a.get("pads", [0, 0, 0, 0])
                pads = (p4[1], p4[3], p4[0], p4[2])
            assert a.get("ceil_mode", 0) == 0
            if n.op == "MaxPool":
                y = F.max_pool2d(F.pad(x, pads, value=float("-inf")), k, st)
            else:
                assert a.get("count_include_pad", 0) == 0 and not any(pads)
                y = F.avg_pool2d(x, k, st)
        elif n.op == "Min":
            y = torch.minimum(i[0], i[1])
        elif n.op == "Less":
            y = i[0] < i[1]
        elif n.op == "Equal":
            y = i[0] == i[1]
        elif n.op == "Not":
            y = ~i[0]
        elif n.op == "And":
            y = i[0] & i[1]
        elif n.op == "Cast":
            y = i[0].to(_TORCH_DT[a["to"]])
        elif n.op == "GlobalAveragePool":
            y = i[0].mean(dim=(2, 3), keepdim=True)
        elif n.op == "Resize":
            assert a.get("mode") == "nearest" and a.get("coordinate_transformation_mode") == "asymmetric" and a.get("nearest_mode", "floor") == "floor"
            ih, iw = i[0].shape[2:]
            if len(i) > 3 and i[3] is not None and i[3].numel() == 4:
                oh, ow = _ints(i[3])[2:]
                sy, sx = oh / ih, ow / iw
            else:
                sc = i[2].numpy()
                sy, sx = float(sc[2]), float(sc[3])
                oh, ow = int(ih * sy), int(iw * sx)
            iy = torch.floor(torch.arange(oh) / sy).long().clamp(max=ih - 1)
            ix = torch.floor(torch.arange(ow) / sx).long().clamp(max=iw - 1)
            y = i[0][:, :, iy][:, :, :, ix]
        elif n.op == "Concat":
            y = torch.cat([t.reshape(-1) if t.dim() == 0 else t for t in i], dim=a.get("axis", 0))
        elif n.op == "Shape":
            y = torch.tensor(list(i[0].shape), dtype=torch.int64)
        elif n.op == "Slice":
            starts, ends = _ints(i[1]), _ints(i[2])
            axes = _ints(i[3]) if len(i) > 3 and i[3] is not None else list(range(len(starts)))
            steps = _ints(i[4]) if len(i) > 4 and i[4] is not None else [1] * len(starts)
            y = i[0]
            for st, en, ax, sp in zip(starts, ends, axes, steps):
                assert sp == 1
                d = y.shape[ax]
                st = max(0, min(d, st + d if st < 0 else st))
                en = max(0, min(d, en + d if en < 0 else en))
                y = y.narrow(ax, st, max(en - st, 0))
        elif n.op == "Squeeze":
            axes = _ints(i[1]) if len(i) > 1 and i[1] is not None else a.get("axes")
            y = i[0]
            if axes is None:
                y = y.squeeze()
            else:
                for ax in sorted(axes, reverse=True):
                    y = y.squeeze(ax)
        elif n.op == "Unsqueeze":
            axes = _ints(i[1]) if len(i) > 1 and i[1] is not None else a.get("axes")
            y = i[0]
            for ax in sorted(axes):
                y = y.unsqueeze(ax)
        elif n.op == "Reshape":
            shp = _ints(i[1])
            shp = [i[0].shape[k] if v == 0 else v for k, v in enumerate(shp)]
            y = i[0].reshape(shp)
        elif n.op == "Transpose":
            y = i[0].permute(a["perm"])
        elif n.op == "ConstantOfShape":
            v = np.asarray(a["value"]).reshape(-1)
            y = torch.full(_ints(i[0]), v[0].item(), dtype=_t(v).dtype)
        elif n.op == "MatMul":
            y = i[0] @ i[1]
        elif n.op == "Softmax":
            y = torch.softmax(i[0], dim=a.get("axis", -1))
        elif n.op == "OneHot":
            depth, vals = int(i[1].reshape(-1)[0]), i[2]
            oh = F.one_hot(i[0].long(), depth)
            y = torch.where(oh.bool(), vals[1], vals[0])
        elif n.op == "Split":
            parts = a.get("split") or (_ints(i[1]) if len(i) > 1 and i[1] is not None else None)
            ax = a.get("axis", 0)
            ys = list(torch.split(i[0], parts if parts else i[0].shape[ax] // len(n.outputs), dim=ax))
        elif n.op == "ArgMax":
            assert a.get("select_last_index", 0) == 0
            y = torch.argmax(i[0], dim=a.get("axis", 0), keepdim=bool(a.get("keepdims", 1)))
        elif n.op == "Gather":
            y = torch.index_select(i[0], a.get("axis", 0), i[1].reshape(-1).long())
            if i[1].dim() == 0:
                y = y.squeeze(a.get("axis", 0))
        elif n.op == "Expand":
            y = i[0].expand(torch.broadcast_shapes(tuple(i[0].shape), tuple(_ints(i[1])))).clone()
        elif n.op == "Range":
            y = torch.arange(int(i[0]), int(i[1]), int(i[2]), dtype=i[0].dtype)
        elif n.op == "Tile":
            y = i[0].repeat(_ints(i[1]))
        elif n.op == "ScatterElements":
            y = i[0].clone().scatter_(a.get("axis", 0), i[1].long(), i[2])
        elif n.op in ("ReduceMax", "ReduceMin"):
            axes = [int(v) - (1 << 64) if int(v) >= (1 << 63) else int(v) for v in a["axes"]]
            y = i[0]
            for ax in axes:
                y = (y.amax if n.op == "ReduceMax" else y.amin)(dim=ax, keepdim=bool(a.get("keepdims", 1)))
        elif n.op == "If":
            br = a["then_branch"] if bool(i[0].reshape(-1)[0]) else a["else_branch"]
            sub = dict(env)
            _exec(br, sub)
            ys = [sub[o] for o in br.outputs]
        elif n.op == "Loop":
            body = a["body"]
            trip = int(i[0].reshape(-1)[0]) if i[0] is not None else (1 << 62)
            cond = bool(i[1].reshape(-1)[0]) if i[1] is not None else True
            carried = list(i[2:])
            it = 0
            while it < trip and cond:
                sub = dict(env)
                sub[body.inputs[0]] = torch.tensor(it, dtype=torch.int64)
                sub[body.inputs[1]] = torch.tensor(cond)
                for name, v in zip(body.inputs[2:], carried):
                    sub[name] = v
                _exec(body, sub)
                cond = bool(sub[body.outputs[0]].reshape(-1)[0])
                carried = [sub[o] for o in body.outputs[1:1 + len(carried)]]
                it += 1
            ys = carried
        else:
            raise NotImplementedError(n.op)
        if ys is None:
            ys = [y]
        for name, v in zip(n.outputs, ys):
            env[name] = v
    return env


def run(path, x, outputs=None):
    """x: numpy [n,3,h,w] float32 -> numpy output of the graph (a list when the graph has several outputs; `outputs` names
    intermediate tensors to return instead)."""
    g = onnx_lite.load(path)
    env = {g.inputs[0]: torch.from_numpy(np.ascontiguousarray(x, np.float32))}
    with torch.no_grad():
        _exec(g, env)
    names = outputs or g.outputs
    res = [env[o].numpy() for o in names]
    return res[0] if len(res) == 1 and outputs is None else res


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
