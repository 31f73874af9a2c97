"""Weight packer: PP-OCRv6-small safetensors -> the flat "RDW1" blob the C-ABI consumes.

All BatchNorms are folded into the preceding conv here (float64 math, stored fp32) and
every tensor is re-laid-out for the NHWC / K-major kernels:
  dense conv   [Cout][KH][KW][Cin]      (reference layout [Cout][Cin][KH][KW])
  depthwise    [KH][KW][C]              (reference [C][1][KH][KW])
  1x1 / linear [N][K]                   (K contiguous = UMMA "K-major" B operand)
  convT 2x2 s2 [dy][dx][Cout][Cin]      (reference [Cin][Cout][2][2])
Reference definitions: rapid_doc/model/ocr/ppocrv6_pytorch/modeling/backbones/rec_lcnetv4.py,
necks/db_fpn.py:288-415, heads/det_db_head.py:52-147, necks/rnn.py:203-379,
heads/rec_multi_head.py:66-77; key prefix handling as rapid_doc/model/ocr/torch.py:102-110.
"""
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHTS_DIR = os.path.join(os.path.dirname(_HERE), "weights")
DET_FILE = "ch_PP-OCRv6_det_small.safetensors"
REC_FILE = "ch_PP-OCRv6_rec_small.safetensors"
DICT_FILE = "ppocrv6_small_dict.txt"

DET_BLOCKS = [  # (cin, cout, stride_h, stride_w, se)  rec_lcnetv4.py:7-23
    [(48, 48, 1, 1, 1), (48, 48, 1, 1, 0)],
    [(48, 96, 2, 2, 0), (96, 96, 1, 1, 1), (96, 96, 1, 1, 0)],
    [(96, 192, 2, 2, 0), (192, 192, 1, 1, 1), (192, 192, 1, 1, 0), (192, 192, 1, 1, 1), (192, 192, 1, 1, 0)],
    [(192, 384, 2, 2, 0), (384, 384, 1, 1, 1), (384, 384, 1, 1, 0)],
]
REC_BLOCKS = [  # rec_lcnetv4.py:26-43
    [(96, 96, 1, 1, 1)],
    [(96, 96, 1, 1, 0), (96, 96, 1, 1, 0)],
    [(96, 192, 2, 1, 0), (192, 192, 1, 1, 1), (192, 192, 1, 1, 0), (192, 192, 1, 1, 1), (192, 192, 1, 1, 0),
     (192, 192, 1, 1, 1), (192, 192, 1, 1, 0)],
    [(192, 384, 2, 1, 0), (384, 384, 1, 1, 1), (384, 384, 1, 1, 0)],
]


def load_safetensors(path):
    from safetensors.numpy import load_file
    sd = load_file(path)
    return {(k[6:] if k.startswith("model.") else k): np.asarray(v, dtype=np.float64) for k, v in sd.items()}


def load_characters(path=None):
    path = path or os.path.join(WEIGHTS_DIR, DICT_FILE)
    return ["blank"] + [l.rstrip("\n") for l in open(path, encoding="utf-8")] + [" "]


def _fold(w, bn, sd, conv_bias=None, eps=1e-5):
    """conv weight [Cout, ...] + BN(prefix bn) -> (w', b')."""
    g, b = sd[bn + ".weight"], sd[bn + ".bias"]
    m, v = sd[bn + ".running_mean"], sd[bn + ".running_var"]
    s = g / np.sqrt(v + eps)
    w2 = w * s.reshape((-1,) + (1,) * (w.ndim - 1))
    b2 = b - m * s
    if conv_bias is not None:
        b2 = b2 + conv_bias * s
    return w2, b2


def _dense(w):   # [Cout,Cin,KH,KW] -> [Cout,KH,KW,Cin]
    return np.ascontiguousarray(w.transpose(0, 2, 3, 1))


def _dw(w):      # [C,1,KH,KW] -> [KH,KW,C]
    return np.ascontiguousarray(w[:, 0].transpose(1, 2, 0))


def _pw(w):      # [N,K,1,1] -> [N,K]
    return np.ascontiguousarray(w.reshape(w.shape[0], w.shape[1]))


def _backbone(sd, blocks, out):
    p = "backbone.encoder.convolution."
    for name in ["stem1", "stem2a", "stem2b", "stem3", "stem4"]:
        w, b = _fold(sd[p + name + ".convolution.weight"], p + name + ".normalization", sd)
        out[name + ".w"] = _dense(w)
        out[name + ".b"] = b
    # stem2a/2b with the intermediate channel count padded to a multiple of 8 (16-byte NHWC pixel pitch
    # for TMA): padded output channels have zero weight + zero bias (relu(0) = 0), padded inputs zero weight
    ch = out["stem2a.w"].shape[0]
    chp = (ch + 7) // 8 * 8
    wa = np.zeros((chp,) + out["stem2a.w"].shape[1:]); wa[:ch] = out["stem2a.w"]
    ba = np.zeros(chp); ba[:ch] = out["stem2a.b"]
    wb = np.zeros(out["stem2b.w"].shape[:3] + (chp,)); wb[..., :ch] = out["stem2b.w"]
    out["stem2a.wp"], out["stem2a.bp"], out["stem2b.wp"] = wa, ba, wb
    for si, stage in enumerate(blocks):
        for bi, (cin, cout, sh, sw, se) in enumerate(stage):
            q = f"backbone.encoder.blocks.{si}.blocks.{bi}."
            n = f"s{si}.b{bi}."
            rep = (sh == 1 and sw == 1 and cin == cout)
            if rep:
                out[n + "dw.w"] = _dw(sd[q + "token_conv.weight"])
                out[n + "dw.b"] = sd[q + "token_conv.bias"]
            else:
                w, b = _fold(sd[q + "token_conv.convolution.weight"], q + "token_conv.normalization", sd)
                out[n + "dw.w"] = _dw(w)
                out[n + "dw.b"] = b
            if se:
                r = q + "token_squeeze_excitation.convolutions."
                out[n + "se.w1"] = _pw(sd[r + "0.weight"]); out[n + "se.b1"] = sd[r + "0.bias"]
                out[n + "se.w2"] = _pw(sd[r + "2.weight"]); out[n + "se.b2"] = sd[r + "2.bias"]
            w, b = _fold(sd[q + "channel_conv1.convolution.weight"], q + "channel_conv1.normalization", sd)
            out[n + "pw1.w"] = _pw(w); out[n + "pw1.b"] = b
            w, b = _fold(sd[q + "channel_conv2.convolution.weight"], q + "channel_conv2.normalization", sd)
            out[n + "pw2.w"] = _pw(w); out[n + "pw2.b"] = b


def det_tensors(path=None):
    sd = load_safetensors(path or os.path.join(WEIGHTS_DIR, DET_FILE))
    out = {}
    _backbone(sd, DET_BLOCKS, out)
    for i in range(4):
        p = f"neck.insert_conv.{i}."
        out[f"neck.in{i}.w"] = _pw(sd[p + "in_conv.weight"])
        q = p + "squeeze_excitation_block."
        out[f"neck.in{i}.se.w1"] = _pw(sd[q + "conv1.weight"]); out[f"neck.in{i}.se.b1"] = sd[q + "conv1.bias"]
        out[f"neck.in{i}.se.w2"] = _pw(sd[q + "conv2.weight"]); out[f"neck.in{i}.se.b2"] = sd[q + "conv2.bias"]
        p = f"neck.input_conv.{i}."
        out[f"neck.lk{i}.dw.w"] = _dw(sd[p + "depthwise_convolution.weight"])
        out[f"neck.lk{i}.dw.b"] = sd[p + "depthwise_convolution.bias"]
        out[f"neck.lk{i}.pw.w"] = _pw(sd[p + "pointwise_convolution.weight"])
        q = p + "squeeze_excitation_module."
        out[f"neck.lk{i}.se.w1"] = _pw(sd[q + "conv1.weight"]); out[f"neck.lk{i}.se.b1"] = sd[q + "conv1.bias"]
        out[f"neck.lk{i}.se.w2"] = _pw(sd[q + "conv2.weight"]); out[f"neck.lk{i}.se.b2"] = sd[q + "conv2.bias"]
    w, b = _fold(sd["head.conv_down.convolution.weight"], "head.conv_down.norm", sd)
    out["head.down.w"] = _dense(w); out["head.down.b"] = b
    # ConvTranspose2d weight [Cin,Cout,2,2]: fold BN over Cout (axis 1)
    wt = sd["head.conv_up.convolution.weight"].transpose(1, 0, 2, 3)          # [Cout,Cin,2,2]
    wt, b = _fold(wt, "head.conv_up.norm", sd, conv_bias=sd["head.conv_up.convolution.bias"])
    out["head.up.w"] = np.ascontiguousarray(wt.transpose(2, 3, 0, 1))          # [dy,dx,Cout,Cin]
    out["head.up.b"] = b
    wf = sd["head.conv_final.weight"]                                          # [Cin,1,2,2]
    out["head.final.w"] = np.ascontiguousarray(wf[:, 0].transpose(1, 2, 0))    # [dy,dx,Cin]
    out["head.final.b"] = sd["head.conv_final.bias"]
    return out


def rec_tensors(path=None):
    sd = load_safetensors(path or os.path.join(WEIGHTS_DIR, REC_FILE))
    out = {}
    _backbone(sd, REC_BLOCKS, out)
    p = "head.encoder."
    for i in range(2):
        w, b = _fold(sd[p + f"conv_block.{i}.convolution.weight"], p + f"conv_block.{i}.normalization", sd)
        out[f"svtr.c{i}.w"] = _pw(w); out[f"svtr.c{i}.b"] = b
    w, b = _fold(sd[p + "conv_block.2.convolution.weight"], p + "conv_block.2.normalization", sd)
    out["svtr.dw.w"] = _dw(w); out["svtr.dw.b"] = b                           # [1,7,120]
    for i in range(2):
        q = p + f"svtr_block.{i}."
        n = f"svtr.blk{i}."
        out[n + "ln1.g"] = sd[q + "layer_norm1.weight"]; out[n + "ln1.b"] = sd[q + "layer_norm1.bias"]
        out[n + "ln2.g"] = sd[q + "layer_norm2.weight"]; out[n + "ln2.b"] = sd[q + "layer_norm2.bias"]
        out[n + "qkv.w"] = sd[q + "self_attn.qkv.weight"]; out[n + "qkv.b"] = sd[q + "self_attn.qkv.bias"]
        out[n + "proj.w"] = sd[q + "self_attn.projection.weight"]; out[n + "proj.b"] = sd[q + "self_attn.projection.bias"]
        out[n + "fc1.w"] = sd[q + "mlp.fc1.weight"]; out[n + "fc1.b"] = sd[q + "mlp.fc1.bias"]
        out[n + "fc2.w"] = sd[q + "mlp.fc2.weight"]; out[n + "fc2.b"] = sd[q + "mlp.fc2.bias"]
    out["svtr.norm.g"] = sd[p + "norm.weight"]; out["svtr.norm.b"] = sd[p + "norm.bias"]
    out["ctc.w"] = sd["head.head.weight"]; out["ctc.b"] = sd["head.head.bias"]
    return out


def pack(tensors) -> bytes:
    """RDW1 blob: magic, n, n * (name[64], ndim, shape[4], offset u64, numel u64), fp32 data
    (each tensor 256-byte aligned)."""
    names = list(tensors)
    hdr = 8 + len(names) * (64 + 4 + 16 + 8 + 8)
    off = (hdr + 255) // 256 * 256
    entries, chunks = [], []
    for n in names:
        a = np.ascontiguousarray(np.asarray(tensors[n], dtype=np.float32))
        shp = list(a.shape) + [1] * (4 - a.ndim)
        entries.append(struct.pack("<64sI4IQQ", n.encode(), a.ndim, *shp, off, a.size))
        chunks.append((off, a.tobytes()))
        off = (off + a.nbytes + 255) // 256 * 256
    buf = bytearray(off)
    buf[0:8] = struct.pack("<4sI", b"RDW1", len(names))
    p = 8
    for e in entries:
        buf[p:p + len(e)] = e
        p += len(e)
    for o, d in chunks:
        buf[o:o + len(d)] = d
    return bytes(buf)


def det_blob(path=None) -> bytes:
    return pack(det_tensors(path))


def rec_blob(path=None) -> bytes:
    return pack(rec_tensors(path))
