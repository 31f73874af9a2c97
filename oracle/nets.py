"""ORACLE (test infrastructure, never on the product path): CPU fp32 restatement of the
PP-OCRv6-small det and rec networks of RapidDoc, written as plain functional torch ops
on the ORIGINAL safetensors weights with BatchNorm UNFOLDED (eps 1e-5), i.e. the same
arithmetic the reference's torch engine runs (rapid_doc/model/ocr/torch.py:171-192).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this file.  It does not need /root/reference: weights come from weights/.

Pinned: tests/test_oracle.py checks this restatement (a) against the reference's own
BaseModel imported from /root/reference when that tree exists (build container) and
(b) against tests/golden/*.npz generated from that import by oracle/make_golden.py.

Reference lines followed (relative to rapid_doc/model/ocr/ppocrv6_pytorch/modeling/):
  stem          backbones/rec_lcnetv4.py:145-169
  SE            backbones/rec_lcnetv4.py:120-142   (nn.Hardsigmoid = clip(x/6+.5,0,1))
  block         backbones/rec_lcnetv4.py:172-236   (residual taken AFTER token mixer + SE)
  det/rec cfg   backbones/rec_lcnetv4.py:7-43, 283-311
  RepLKFPN      necks/db_fpn.py:288-415            (SE = clip(.2x+.5,0,1), residual form)
  DBHead v6     heads/det_db_head.py:52-147
  LightSVTR     necks/rnn.py:203-379
  MultiHead     heads/rec_multi_head.py:66-77      (raw logits; engine softmaxes)
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHTS_DIR = os.path.join(os.path.dirname(_HERE), "weights")

DET_BLOCKS = [  # (k, cin, cout, stride, se)   rec_lcnetv4.py:7-23
    [(3, 48, 48, 1, True), (3, 48, 48, 1, False)],
    [(3, 48, 96, 2, False), (3, 96, 96, 1, True), (3, 96, 96, 1, False)],
    [(3, 96, 192, 2, False), (3, 192, 192, 1, True), (3, 192, 192, 1, False), (3, 192, 192, 1, True), (3, 192, 192, 1, False)],
    [(3, 192, 384, 2, False), (3, 384, 384, 1, True), (3, 384, 384, 1, False)],
]
REC_BLOCKS = [  # rec_lcnetv4.py:26-43
    [(3, 96, 96, 1, True)],
    [(3, 96, 96, 1, False), (3, 96, 96, 1, False)],
    [(3, 96, 192, (2, 1), False), (3, 192, 192, 1, True), (3, 192, 192, 1, False), (3, 192, 192, 1, True),
     (3, 192, 192, 1, False), (3, 192, 192, 1, True), (3, 192, 192, 1, False)],
    [(3, 192, 384, (2, 1), False), (3, 384, 384, 1, True), (3, 384, 384, 1, False)],
]


def load_state(name):
    from safetensors.torch import load_file
    sd = load_file(os.path.join(WEIGHTS_DIR, name))
    return {k.removeprefix("model."): v.float() for k, v in sd.items()}  # torch.py:102-110


def load_characters():
    """['blank'] + dict lines + [' ']  (CTCLabelDecode)."""
    path = os.path.join(WEIGHTS_DIR, "ppocrv6_small_dict.txt")
    chars = ["blank"] + [l.rstrip("\n") for l in open(path, encoding="utf-8")] + [" "]
    return chars


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _cba(x, sd, p, stride=1, groups=1, act=None, k=None):
    """PPLCNetV4ConvLayer: conv(no bias, pad (k-1)//2) + BN + act.  rec_lcnetv4.py:86-117"""
    w = sd[p + ".convolution.weight"]
    k = w.shape[-1]
    x = F.conv2d(x, w, None, stride, (k - 1) // 2, 1, groups)
    x = _bn(x, sd, p + ".normalization")
    if act == "relu":
        x = F.relu(x)
    return x


def stem(x, sd, p="backbone.encoder.convolution"):
    e = _cba(x, sd, p + ".stem1", stride=2, act="relu")
    e = F.pad(e, (0, 1, 0, 1))
    a = _cba(e, sd, p + ".stem2a", act="relu")
    a = F.pad(a, (0, 1, 0, 1))
    a = _cba(a, sd, p + ".stem2b", act="relu")
    pooled = F.max_pool2d(e, kernel_size=2, stride=1, ceil_mode=True)
    e = torch.cat([pooled, a], dim=1)
    e = _cba(e, sd, p + ".stem3", stride=2, act="relu")
    e = _cba(e, sd, p + ".stem4", act="relu")
    return e


def lcnet_block(x, sd, p, k, cin, cout, stride, se):
    rep = (stride == 1) and (cin == cout)
    if rep:  # plain dw conv with bias
        x = F.conv2d(x, sd[p + ".token_conv.weight"], sd[p + ".token_conv.bias"], 1, k // 2, 1, cin)
    else:    # dw conv + BN, no act
        x = _cba(x, sd, p + ".token_conv", stride=stride, groups=cin)
    if se:
        q = p + ".token_squeeze_excitation.convolutions"
        s = F.adaptive_avg_pool2d(x, 1)
        s = F.relu(F.conv2d(s, sd[q + ".0.weight"], sd[q + ".0.bias"]))
        s = F.hardsigmoid(F.conv2d(s, sd[q + ".2.weight"], sd[q + ".2.bias"]))
        x = x * s
    res = x
    x = _cba(x, sd, p + ".channel_conv1")
    x = F.gelu(x)
    x = _cba(x, sd, p + ".channel_conv2")
    if rep:
        x = res + x
    return x


def backbone(x, sd, cfg):
    x = stem(x, sd)
    feats = []
    for si, stage in enumerate(cfg):
        for bi, (k, cin, cout, stride, se) in enumerate(stage):
            x = lcnet_block(x, sd, f"backbone.encoder.blocks.{si}.blocks.{bi}", k, cin, cout, stride, se)
        feats.append(x)
    return feats


def _fpn_se(x, sd, p):
    """RepLKFPNSqueezeExcitationModule  db_fpn.py:288-308  -> returns x * gate."""
    s = F.adaptive_avg_pool2d(x, 1)
    s = F.conv2d(F.relu(F.conv2d(s, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"])
    s = torch.clamp(0.2 * s + 0.5, min=0.0, max=1.0)
    return x * s


def rep_lk_fpn(feats, sd):
    fused = []
    for i, f in enumerate(feats):
        h = F.conv2d(f, sd[f"neck.insert_conv.{i}.in_conv.weight"])
        fused.append(h + _fpn_se(h, sd, f"neck.insert_conv.{i}.squeeze_excitation_block"))
    for i in range(2, -1, -1):
        fused[i] = fused[i] + F.interpolate(fused[i + 1], scale_factor=2, mode="nearest")
    outs = []
    for i, f in enumerate(fused):
        p = f"neck.input_conv.{i}"
        h = F.conv2d(f, sd[p + ".depthwise_convolution.weight"], sd[p + ".depthwise_convolution.bias"], 1, 3, 1, 96)
        h = F.conv2d(h, sd[p + ".pointwise_convolution.weight"])
        h = h + _fpn_se(h, sd, p + ".squeeze_excitation_module")
        outs.append(h)
    proc = [outs[0]] + [F.interpolate(o, scale_factor=s, mode="nearest") for o, s in zip(outs[1:], [2, 4, 8])]
    return torch.cat(proc[::-1], dim=1)


def db_head(x, sd):
    x = F.conv2d(x, sd["head.conv_down.convolution.weight"], None, 1, 1)
    x = F.relu(_bn(x, sd, "head.conv_down.norm"))
    x = F.conv_transpose2d(x, sd["head.conv_up.convolution.weight"], sd["head.conv_up.convolution.bias"], 2)
    x = F.relu(_bn(x, sd, "head.conv_up.norm"))
    x = F.conv_transpose2d(x, sd["head.conv_final.weight"], sd["head.conv_final.bias"], 2)
    return torch.nan_to_num(torch.sigmoid(x))


_DET_SD = None
_REC_SD = None


def det_state():
    global _DET_SD
    if _DET_SD is None:
        _DET_SD = load_state("ch_PP-OCRv6_det_small.safetensors")
    return _DET_SD


def rec_state():
    global _REC_SD
    if _REC_SD is None:
        _REC_SD = load_state("ch_PP-OCRv6_rec_small.safetensors")
    return _REC_SD


@torch.no_grad()
def det_forward(x: np.ndarray, return_feats=False):
    """[B,3,H,W] f32 (H,W multiples of 32) -> prob map [B,1,H,W] f32."""
    sd = det_state()
    x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    feats = backbone(x, sd, DET_BLOCKS)
    neck = rep_lk_fpn(feats, sd)
    out = db_head(neck, sd)
    if return_feats:
        return out.numpy(), [f.numpy() for f in feats], neck.numpy()
    return out.numpy()


def _svtr_conv(x, sd, p, pad=(0, 0), groups=1):
    x = F.conv2d(x, sd[p + ".convolution.weight"], None, 1, pad, 1, groups)
    return F.silu(_bn(x, sd, p + ".normalization"))


def light_svtr(x, sd, p="head.encoder", heads=8):
    res = _svtr_conv(x, sd, p + ".conv_block.0")
    h = _svtr_conv(x, sd, p + ".conv_block.1")
    h = h + _svtr_conv(h, sd, p + ".conv_block.2", pad=(0, 3), groups=h.shape[1])
    B, C, H, W = h.shape
    t = h.flatten(2).permute(0, 2, 1)
    for i in range(2):
        q = f"{p}.svtr_block.{i}"
        r = t
        y = F.layer_norm(t, (C,), sd[q + ".layer_norm1.weight"], sd[q + ".layer_norm1.bias"], 1e-6)
        qkv = F.linear(y, sd[q + ".self_attn.qkv.weight"], sd[q + ".self_attn.qkv.bias"])
        T = qkv.shape[1]
        qkv = qkv.reshape(B, T, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
        att = torch.matmul(qkv[0], qkv[1].transpose(-1, -2)) * ((C // heads) ** -0.5)
        att = F.softmax(att, dim=-1)
        y = torch.matmul(att, qkv[2]).transpose(1, 2).reshape(B, T, C)
        t = r + F.linear(y, sd[q + ".self_attn.projection.weight"], sd[q + ".self_attn.projection.bias"])
        r = t
        y = F.layer_norm(t, (C,), sd[q + ".layer_norm2.weight"], sd[q + ".layer_norm2.bias"], 1e-6)
        y = F.silu(F.linear(y, sd[q + ".mlp.fc1.weight"], sd[q + ".mlp.fc1.bias"]))
        t = r + F.linear(y, sd[q + ".mlp.fc2.weight"], sd[q + ".mlp.fc2.bias"])
    t = F.layer_norm(t, (C,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    return t.reshape(B, H, W, C).permute(0, 3, 1, 2) + res


@torch.no_grad()
def rec_logits(x: np.ndarray, return_feats=False):
    """[B,3,48,W] f32 -> raw CTC logits [B,W/8,18710] f32 (MultiHead.forward eval)."""
    sd = rec_state()
    x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    f = backbone(x, sd, REC_BLOCKS)[-1]
    f = F.avg_pool2d(f, [3, 2])                       # rec_lcnetv4.py:309-311
    e = light_svtr(f, sd)
    seq = e.squeeze(2).permute(0, 2, 1)
    logits = F.linear(seq, sd["head.head.weight"], sd["head.head.bias"])
    if return_feats:
        return logits.numpy(), f.numpy(), seq.numpy()
    return logits.numpy()


@torch.no_grad()
def rec_forward(x: np.ndarray):
    """Engine-level output the reference hands to CTCLabelDecode: softmax(logits, dim=2)
    (rapid_doc/model/ocr/torch.py:186-187)."""
    return torch.softmax(torch.from_numpy(rec_logits(x)), dim=2).numpy()
