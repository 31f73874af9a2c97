"""CPU tests of the weight packer: BN folding is exact (float64) and the RDW1 blob round-trips."""
import struct

import numpy as np
import torch
import torch.nn.functional as F

from oracle import nets
from rapiddoc_b200 import weights as W


def _unpack(blob):
    assert blob[:4] == b"RDW1"
    (n,) = struct.unpack_from("<I", blob, 4)
    out = {}
    for i in range(n):
        name, nd, s0, s1, s2, s3, off, numel = struct.unpack_from("<64sI4IQQ", blob, 8 + i * 100)
        shape = (s0, s1, s2, s3)[:nd]
        out[name.rstrip(b"\0").decode()] = np.frombuffer(blob, np.float32, numel, off).reshape(shape)
    return out


def test_pack_roundtrip_and_alignment():
    t = {"a.w": np.arange(24, dtype=np.float64).reshape(2, 3, 4), "b": np.ones(5)}
    blob = W.pack(t)
    u = _unpack(blob)
    assert np.array_equal(u["a.w"], t["a.w"].astype(np.float32)) and u["b"].shape == (5,)
    for i in range(2):
        off = struct.unpack_from("<Q", blob, 8 + i * 100 + 84)[0]
        assert off % 256 == 0


def test_bn_folding_matches_unfolded_conv_bn():
    sd = nets.det_state()
    t = W.det_tensors()
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((1, 3, 32, 32)).astype(np.float32))
    p = "backbone.encoder.convolution.stem1"
    want = nets._cba(x, sd, p, stride=2)                                   # conv + BN, unfolded
    w = torch.from_numpy(t["stem1.w"].astype(np.float32)).permute(0, 3, 1, 2)   # [Cout,KH,KW,Cin] -> OIHW
    got = F.conv2d(x, w, torch.from_numpy(t["stem1.b"].astype(np.float32)), 2, 1)
    assert (got - want).abs().max() < 2e-5
    # padded stem2a/2b variants: extra output channels are exactly zero, extra input channels have zero weight
    assert np.all(t["stem2a.wp"][12:] == 0) and np.all(t["stem2a.bp"][12:] == 0) and np.all(t["stem2b.wp"][..., 12:] == 0)
    # ConvTranspose + BN of the DB head
    xh = torch.from_numpy(np.random.default_rng(1).standard_normal((1, 24, 5, 7)).astype(np.float32))
    want = nets._bn(F.conv_transpose2d(xh, sd["head.conv_up.convolution.weight"], sd["head.conv_up.convolution.bias"], 2), sd, "head.conv_up.norm")
    wu = torch.from_numpy(t["head.up.w"].astype(np.float32))               # [dy,dx,Cout,Cin]
    got = torch.zeros_like(want)
    for dy in range(2):
        for dx in range(2):
            got[:, :, dy::2, dx::2] = torch.einsum("oc,nchw->nohw", wu[dy, dx], xh) + torch.from_numpy(t["head.up.b"].astype(np.float32))[None, :, None, None]
    assert (got - want).abs().max() < 2e-5


def test_all_tensors_present_and_finite():
    for tensors, must in ((W.det_tensors(), ["stem1.w", "s3.b2.pw2.w", "neck.in3.w", "neck.lk0.dw.w", "head.final.w"]),
                          (W.rec_tensors(), ["stem1.w", "s2.b6.pw2.w", "svtr.blk1.fc2.w", "svtr.dw.w", "ctc.w"])):
        for m in must:
            assert m in tensors
        assert all(np.isfinite(v).all() for v in tensors.values())
    assert W.rec_tensors()["ctc.w"].shape == (18710, 120) and len(W.load_characters()) == 18710
