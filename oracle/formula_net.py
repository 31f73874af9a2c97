"""ORACLE (test infrastructure, never on the product path): PP-FormulaNet_plus on CPU, plain functional torch fp32, BatchNorm
UNFOLDED, on a state_dict in the reference's key layout.  Restates
  rapid_doc/model/formula/rapid_formula_self/networks/backbones/rec_pphgnetv2.py:858-1207,1587-1642  (PPHGNetV2-B6 encoder)
  .../networks/heads/rec_unimernet_head.py:440-456,502-748,931-976                                   (MBart decoder layers)
  .../networks/heads/rec_ppformulanet_head.py:400-632,1052-1171                                      (generate_export, greedy)
Pinned: tests/test_formula.py runs it against the reference's own `BaseModel` imported from /root/reference (build container)
on the same seeded weights (ids identical, encoder output to 1e-4), and against tests/golden/formula_m_ids.npz made from that
import.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this file."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _cba(x, sd, p, stride=1, padding=0, groups=1, act=True):
    w = sd[p + ".conv.weight"]
    if padding == "same2":                       # 2x2 conv, padding='same': one zero row / column at the bottom / right
        x = F.pad(x, (0, 1, 0, 1))
        padding = 0
    x = F.conv2d(x, w, None, stride, padding, 1, groups)
    x = F.batch_norm(x, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"], sd[p + ".bn.bias"], False, 0.0, 1e-5)
    return F.relu(x) if act else x


def encoder(x, sd, arch):
    p = "backbone.pphgnet_b6."
    if x.shape[1] == 1:
        x = torch.repeat_interleave(x, repeats=3, dim=1)
    x = _cba(x, sd, p + "stem.stem1", 2, 1)
    x2 = _cba(_cba(x, sd, p + "stem.stem2a", 1, "same2"), sd, p + "stem.stem2b", 1, "same2")
    x1 = F.max_pool2d(F.pad(x, (0, 1, 0, 1)), 2, 1, 0, ceil_mode=True)
    x = torch.cat([x1, x2], 1)
    x = _cba(x, sd, p + "stem.stem3", 2, 1)
    x = _cba(x, sd, p + "stem.stem4", 1, 0)
    for si, (cin, mid, cout, blocks, down, light, k, layers) in enumerate(arch["stages"]):
        sp = f"{p}stages.{si}."
        if down:
            x = _cba(x, sd, sp + "downsample", 2, 1, cin, act=False)
        for b in range(blocks):
            bp = f"{sp}blocks.{b}."
            ident, outs = x, [x]
            for l in range(layers):
                if light:
                    x = _cba(x, sd, f"{bp}layers.{l}.conv1", 1, 0, act=False)
                    x = _cba(x, sd, f"{bp}layers.{l}.conv2", 1, (k - 1) // 2, mid)
                else:
                    x = _cba(x, sd, f"{bp}layers.{l}", 1, (k - 1) // 2)
                outs.append(x)
            x = _cba(torch.cat(outs, 1), sd, bp + "aggregation_squeeze_conv")
            x = _cba(x, sd, bp + "aggregation_excitation_conv")
            if b > 0:
                x = x + ident
    b, c, h, w = x.shape
    return x.reshape(b, c, h * w).permute(0, 2, 1)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(x, sd, p):
    return F.layer_norm(x, [x.shape[-1]], sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def generate(enc, sd, arch, max_new_tokens):
    dp = "head.decoder.model.decoder."
    d, H = arch["d_model"], arch["heads"]
    hd = d // H
    B = enc.shape[0]
    ench = _lin(enc, sd, "head.enc_to_dec_proj")
    ids = torch.full((B, 1), arch["start"], dtype=torch.int64)
    unfinished = torch.ones(B, dtype=torch.int64)
    kc = [None] * arch["layers"]
    vc = [None] * arch["layers"]
    cross = []
    for l in range(arch["layers"]):
        lp = f"{dp}layers.{l}.encoder_attn."
        cross.append((_lin(ench, sd, lp + "k_proj").reshape(B, -1, H, hd).permute(0, 2, 1, 3), _lin(ench, sd, lp + "v_proj").reshape(B, -1, H, hd).permute(0, 2, 1, 3)))
    cur = ids
    for step in range(max_new_tokens):
        x = F.embedding(cur, sd[dp + "embed_tokens.weight"]) * math.sqrt(d) + sd[dp + "embed_positions.weight"][step + 2][None, None]
        h = _ln(x, sd, dp + "layernorm_embedding")
        for l in range(arch["layers"]):
            lp = f"{dp}layers.{l}."
            r = h
            x = _ln(h, sd, lp + "self_attn_layer_norm")
            q = (_lin(x, sd, lp + "self_attn.q_proj") * hd ** -0.5).reshape(B, 1, H, hd).permute(0, 2, 1, 3)
            k = _lin(x, sd, lp + "self_attn.k_proj").reshape(B, 1, H, hd).permute(0, 2, 1, 3)
            v = _lin(x, sd, lp + "self_attn.v_proj").reshape(B, 1, H, hd).permute(0, 2, 1, 3)
            kc[l] = k if kc[l] is None else torch.cat([kc[l], k], 2)
            vc[l] = v if vc[l] is None else torch.cat([vc[l], v], 2)
            a = torch.softmax(q @ kc[l].transpose(2, 3), -1) @ vc[l]
            h = r + _lin(a.permute(0, 2, 1, 3).reshape(B, 1, d), sd, lp + "self_attn.out_proj")
            r = h
            x = _ln(h, sd, lp + "encoder_attn_layer_norm")
            q = (_lin(x, sd, lp + "encoder_attn.q_proj") * hd ** -0.5).reshape(B, 1, H, hd).permute(0, 2, 1, 3)
            a = torch.softmax(q @ cross[l][0].transpose(2, 3), -1) @ cross[l][1]
            h = r + _lin(a.permute(0, 2, 1, 3).reshape(B, 1, d), sd, lp + "encoder_attn.out_proj")
            r = h
            x = _ln(h, sd, lp + "final_layer_norm")
            h = r + _lin(F.gelu(_lin(x, sd, lp + "fc1")), sd, lp + "fc2")
        h = _ln(h, sd, dp + "layer_norm")
        logits = F.linear(h[:, -1], sd["head.decoder.lm_head.weight"])
        if ids.shape[-1] == arch["forced_eos_len"] - 1:
            logits = torch.full_like(logits, -math.inf)
            logits[:, arch["eos"]] = 0
        nxt = torch.argmax(logits, -1)
        nxt = nxt * unfinished + arch["pad"] * (1 - unfinished)
        ids = torch.cat([ids, nxt[:, None]], -1)
        cur = nxt[:, None]
        unfinished = unfinished & ~(nxt == arch["eos"]).to(torch.int64)
        if ((ids == arch["eos"]).sum(1) >= 1).all():
            break
    return ids


def forward(x, sd, arch, max_new_tokens):
    """x [B,1,H,W] float32 numpy -> (ids [B,L] int64 numpy, encoder output [B,S,E] float32 numpy)."""
    with torch.no_grad():
        sd = {k: v.float() for k, v in sd.items()}
        enc = encoder(torch.from_numpy(np.asarray(x, np.float32)), sd, arch)
        return generate(enc, sd, arch, max_new_tokens).numpy(), enc.numpy()
