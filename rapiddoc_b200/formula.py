"""Formula path: PP-FormulaNet_plus on the B200 — PPHGNetV2-B6 encoder + MBart decoder with greedy
generation, behind the reference's facade / session / plugin surfaces.

Reference (under rapid_doc/model/formula/):
  F1  RapidFormulaModel.batch_predict            rapid_formula_model.py:34-41 -> rapid_formula_self/main.py:28-41 (chunks of batch_size)
  F2  PPPreProcess                               rapid_formula_self/model_handler/pp_formulanet_plus/pre_process.py:12-257
  F3  PPHGNetV2_B6_Formula                       rapid_formula_self/networks/backbones/rec_pphgnetv2.py:858-1207,1587-1642
  F4  PPFormulaNet_Head (MBart, generate_export) networks/heads/rec_ppformulanet_head.py:400-632,695-803,1052-1171;
                                                 MBart layers networks/heads/rec_unimernet_head.py:440-456,502-748,931-976
  F5  UniMERNetDecode                            model_handler/pp_formulanet_plus/post_process.py:29-408 (tokenizer + LaTeX normalisation)

The trained weights (PP-FormulaNet_plus-M .pth, tokenizer json) are downloaded by RapidDoc at first run and are NOT available
offline: `FormulaEngine` takes a state_dict in the reference's own key layout (`backbone.pphgnet_b6...`, `head.decoder...`),
so the released checkpoint drops in; tests and the bench use seeded synthetic weights of the exact architecture.

Engine design (device side = the `rdb_op_*` entry points of the C-ABI):
  * BatchNorm is folded into the conv weights at load (float64); conv weights are packed [O][kh][kw][C] = the K order of the
    one-pass im2col, so every dense conv is im2col + GEMM and every 1x1 conv / Linear is a GEMM straight on the activations;
    RDB_PREC_FP32 -> fp32 SIMT GEMM (exact-parity mode), RDB_PREC_FP16 -> fp16 activations + tcgen05/TMEM/TMA GEMM.
  * The dense concatenation of an HGV2 block (input + 6 layer outputs, up to 6656 channels) is never materialised by a copy:
    each layer writes its channel slice of the block's wide buffer and reads its input slice in place (row pitch = total).
  * Decoder: KV cache resident on the device, one new token per step for the whole batch; the k/v projections write
    straight into their cache rows; cross-attention K/V are computed once per batch; lm-head GEMM -> row argmax ->
    greedy bookkeeping (eos / pad / forced eos) all on the device, one tiny D2H every `sync_every` steps to test "all done".
"""
import math

import numpy as np

from . import _lib

# stage table of PPHGNetV2_B6 (rec_pphgnetv2.py:1601-1607): in, mid, out, blocks, downsample, light_block, kernel, layers
STAGES_B6 = [(96, 96, 192, 2, False, False, 3, 6), (192, 192, 512, 3, True, False, 3, 6), (512, 384, 1024, 6, True, True, 5, 6),
             (1024, 768, 2048, 3, True, True, 5, 6)]
ARCH_M = dict(stem=(3, 48, 96), stages=STAGES_B6, enc_dim=2048, d_model=512, heads=16, ffn=2048, layers=6, vocab=50000,
              max_new_tokens=2560, forced_eos_len=1537, eos=2, pad=1, start=0, input_size=(384, 384))
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


# ------------------------------------------------------------------------------------------------ F2 preprocessing
class FormulaPreProcess:
    """crop_margin -> short side to min(input_size) (PIL bilinear) -> thumbnail -> centre pad (black) ->
    (x/255 - 0.7931)/0.1738 -> gray -> pad to a multiple of 16 with 1.0 -> [1,1,H,W] float32."""

    def __init__(self, img_size=(384, 384)):
        self.input_size = img_size

    @staticmethod
    def crop_margin(img):
        import cv2
        data = np.array(img.convert("L")).astype(np.uint8)
        mx, mn = data.max(), data.min()
        if mx == mn:
            return img
        data = (data - mn) / (mx - mn) * 255
        gray = 255 * (data < 200).astype(np.uint8)
        a, b, w, h = cv2.boundingRect(cv2.findNonZero(gray))
        return img.crop((a, b, w + a, h + b))

    def decode(self, img):
        from PIL import Image, ImageOps
        try:
            img = self.crop_margin(Image.fromarray(img).convert("RGB"))
        except OSError:
            return None
        if img.height == 0 or img.width == 0:
            return None
        w, h = img.size
        short, long = (w, h) if w <= h else (h, w)
        ns = min(self.input_size)
        nl = int(ns * long / short)
        nw, nh = (ns, nl) if w <= h else (nl, ns)
        img = img.resize((nw, nh), resample=2)
        img.thumbnail((self.input_size[1], self.input_size[0]))
        dw, dh = self.input_size[1] - img.width, self.input_size[0] - img.height
        pw, ph = dw // 2, dh // 2
        return np.array(ImageOps.expand(img, (pw, ph, dw - pw, dh - ph)))

    @staticmethod
    def transform(img):
        import cv2
        mean = np.array([0.7931] * 3).reshape((1, 1, 3)).astype("float32")
        std = np.array([0.1738] * 3).reshape((1, 1, 3)).astype("float32")
        x = (img.astype("float32") * float(1 / 255.0) - mean) / std
        g = np.squeeze(cv2.cvtColor(x, cv2.COLOR_BGR2GRAY))
        return cv2.merge([g] * 3)

    @staticmethod
    def fmt(img):
        h, w = img.shape[:2]
        x = np.pad(img[:, :, 0], ((0, math.ceil(h / 16) * 16 - h), (0, math.ceil(w / 16) * 16 - w)), constant_values=(1, 1))
        return x[:, :, np.newaxis].transpose(2, 0, 1)[np.newaxis, :]

    def __call__(self, imgs):
        return [self.fmt(self.transform(self.decode(im))) for im in imgs]


# ------------------------------------------------------------------------------------------------ synthetic weights
def synthetic_state_dict(arch=ARCH_M, seed=0):
    """Seeded weights in the REFERENCE's state_dict layout, with statistics that keep the signal alive through ~50 conv layers
    (so that outputs depend on the input and greedy decoding is not degenerate).  CPU torch generator: identical on every box."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv_bn(name, cin, cout, k, groups=1):
        fan = (cin // groups) * k * k
        sd[name + ".conv.weight"] = torch.randn(cout, cin // groups, k, k, generator=g) * math.sqrt(2.0 / fan)
        sd[name + ".bn.weight"] = torch.rand(cout, generator=g) * 0.4 + 0.6       # mean 0.8: keeps |activation| O(1..10) through the net
        sd[name + ".bn.bias"] = torch.randn(cout, generator=g) * 0.1
        sd[name + ".bn.running_mean"] = torch.randn(cout, generator=g) * 0.1
        sd[name + ".bn.running_var"] = torch.rand(cout, generator=g) * 0.5 + 0.75

    p = "backbone.pphgnet_b6."
    c0, c1, c2 = arch["stem"]
    conv_bn(p + "stem.stem1", c0, c1, 3)
    conv_bn(p + "stem.stem2a", c1, c1 // 2, 2)
    conv_bn(p + "stem.stem2b", c1 // 2, c1, 2)
    conv_bn(p + "stem.stem3", c1 * 2, c1, 3)
    conv_bn(p + "stem.stem4", c1, c2, 1)
    for si, (cin, mid, cout, blocks, down, light, k, layers) in enumerate(arch["stages"]):
        sp = f"{p}stages.{si}."
        if down:
            conv_bn(sp + "downsample", cin, cin, 3, groups=cin)
        for b in range(blocks):
            bp = f"{sp}blocks.{b}."
            ic = cin if b == 0 else cout
            for l in range(layers):
                lc = ic if l == 0 else mid
                if light:
                    conv_bn(f"{bp}layers.{l}.conv1", lc, mid, 1)
                    conv_bn(f"{bp}layers.{l}.conv2", mid, mid, k, groups=mid)
                else:
                    conv_bn(f"{bp}layers.{l}", lc, mid, k)
            conv_bn(bp + "aggregation_squeeze_conv", ic + layers * mid, cout // 2, 1)
            conv_bn(bp + "aggregation_excitation_conv", cout // 2, cout, 1)
    d, ffn, V = arch["d_model"], arch["ffn"], arch["vocab"]

    def lin(name, i, o, bias=True, std=None):
        sd[name + ".weight"] = torch.randn(o, i, generator=g) * (std if std else 1.0 / math.sqrt(i))
        if bias:
            sd[name + ".bias"] = torch.randn(o, generator=g) * 0.02

    def ln(name, n):
        sd[name + ".weight"] = torch.rand(n, generator=g) * 0.5 + 0.75
        sd[name + ".bias"] = torch.randn(n, generator=g) * 0.05

    lin("head.enc_to_dec_proj", arch["enc_dim"], d)
    dp = "head.decoder.model.decoder."
    sd[dp + "embed_tokens.weight"] = torch.randn(V, d, generator=g) * 0.05
    sd[dp + "embed_positions.weight"] = torch.randn(arch["max_new_tokens"] + 2, d, generator=g) * 0.5
    for l in range(arch["layers"]):
        lp = f"{dp}layers.{l}."
        for a in ("self_attn", "encoder_attn"):
            for q in ("k_proj", "v_proj", "q_proj", "out_proj"):
                lin(f"{lp}{a}.{q}", d, d)
            ln(f"{lp}{a}_layer_norm", d)
        lin(lp + "fc1", d, ffn)
        lin(lp + "fc2", ffn, d)
        ln(lp + "final_layer_norm", d)
    ln(dp + "layernorm_embedding", d)
    ln(dp + "layer_norm", d)
    lin("head.decoder.lm_head", d, V, bias=False, std=0.3)
    return sd


# ------------------------------------------------------------------------------------------------ engine
class _Conv:
    __slots__ = ("w", "w16", "b", "k", "cin", "cout", "dw")


class FormulaEngine:
    """state_dict (reference key layout) -> token ids.  x: [B,1,H,W] float32 (numpy or device tensor), H = W = 384 for -M."""

    def __init__(self, state_dict, device=0, precision=_lib.PREC_FP32, arch=ARCH_M, max_new_tokens=None, sync_every=8, use_graph=True, implicit_conv=True):
        import torch
        self.lib = _lib.load()
        if self.lib.rdb_device_count() <= int(device):
            raise _lib.B200Error("no CUDA device visible — the formula engine has no CPU fallback")
        self.torch = torch
        self.device, self.prec, self.arch = int(device), int(precision), dict(arch)
        self.dev = torch.device("cuda", self.device)
        self.max_new = int(max_new_tokens if max_new_tokens is not None else arch["max_new_tokens"])
        self.sync_every = sync_every
        self.use_graph = use_graph
        self.implicit_conv = implicit_conv       # fp16: dense k x k convs through rdb_op_conv_tc instead of im2col + GEMM
        import os
        self.qkv_parallel = os.environ.get("RDB_FORMULA_QKV", "parallel") != "serial"
        with torch.cuda.device(self.dev):
            self._side = [torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)]
        self._dec = {}
        self.launches = 0
        self.adt = torch.float16 if self.prec == _lib.PREC_FP16 else torch.float32
        self._pack(state_dict)

    # ---------------------------------------------------------------- weights
    def _fold(self, sd, name, depthwise=False, pad_cin=0):
        torch = self.torch
        w = sd[name + ".conv.weight"].double()
        if pad_cin and w.shape[1] < pad_cin:            # zero input channels: keeps GEMM rows 16-byte aligned (K = kh*kw*cin)
            w = torch.cat([w, torch.zeros(w.shape[0], pad_cin - w.shape[1], w.shape[2], w.shape[3], dtype=w.dtype)], 1)
        g, b = sd[name + ".bn.weight"].double(), sd[name + ".bn.bias"].double()
        m, v = sd[name + ".bn.running_mean"].double(), sd[name + ".bn.running_var"].double()
        s = g / torch.sqrt(v + 1e-5)
        w = w * s[:, None, None, None]
        c = _Conv()
        c.cout, c.k, c.dw = w.shape[0], w.shape[2], depthwise
        c.b = (b - m * s).float().to(self.dev).contiguous()
        if depthwise:                                   # [C,1,k,k] -> [k][k][C]
            c.cin = c.cout
            c.w = w[:, 0].permute(1, 2, 0).float().to(self.dev).contiguous()
            c.w16 = None
        else:                                           # [O,C,kh,kw] -> [O][kh][kw][C]
            c.cin = w.shape[1]
            c.w = w.permute(0, 2, 3, 1).reshape(c.cout, -1).float().to(self.dev).contiguous()
            c.w16 = c.w.half().contiguous() if self.prec == _lib.PREC_FP16 else None
        return c

    def _lin(self, sd, name, scale=None):
        w = sd[name + ".weight"].double()
        b = sd.get(name + ".bias")
        if scale is not None:
            w = w * scale
            b = b.double() * scale if b is not None else None
        return (w.float().to(self.dev).contiguous(), None if b is None else b.float().to(self.dev).contiguous())

    def _pack(self, sd):
        a = self.arch
        p = "backbone.pphgnet_b6."
        self.stem = {k: self._fold(sd, f"{p}stem.{k}", pad_cin=4 if k == "stem1" else 0) for k in ("stem1", "stem2a", "stem2b", "stem3", "stem4")}
        self.stages = []
        for si, (cin, mid, cout, blocks, down, light, k, layers) in enumerate(a["stages"]):
            sp = f"{p}stages.{si}."
            st = {"down": self._fold(sd, sp + "downsample", True) if down else None, "blocks": []}
            for b in range(blocks):
                bp = f"{sp}blocks.{b}."
                ls = []
                for l in range(layers):
                    if light:
                        ls.append((self._fold(sd, f"{bp}layers.{l}.conv1"), self._fold(sd, f"{bp}layers.{l}.conv2", True)))
                    else:
                        ls.append((self._fold(sd, f"{bp}layers.{l}"), None))
                st["blocks"].append({"layers": ls, "sq": self._fold(sd, bp + "aggregation_squeeze_conv"),
                                     "ex": self._fold(sd, bp + "aggregation_excitation_conv"), "identity": b > 0})
            self.stages.append(st)
        d, H = a["d_model"], a["heads"]
        scaling = (d // H) ** -0.5
        dp = "head.decoder.model.decoder."
        self.proj = self._lin(sd, "head.enc_to_dec_proj")
        self.tok = sd[dp + "embed_tokens.weight"].float().to(self.dev).contiguous()
        self.pos = sd[dp + "embed_positions.weight"].float().to(self.dev).contiguous()

        def ln(name):
            return (sd[name + ".weight"].float().to(self.dev).contiguous(), sd[name + ".bias"].float().to(self.dev).contiguous())
        self.ln_emb, self.ln_out = ln(dp + "layernorm_embedding"), ln(dp + "layer_norm")
        self.layers = []
        for l in range(a["layers"]):
            lp = f"{dp}layers.{l}."
            L = {}
            for att, key in (("self_attn", "s"), ("encoder_attn", "c")):
                L[key + "q"] = self._lin(sd, f"{lp}{att}.q_proj", scaling)       # q = (x W^T + b) * head_dim^-0.5, folded
                L[key + "k"] = self._lin(sd, f"{lp}{att}.k_proj")
                L[key + "v"] = self._lin(sd, f"{lp}{att}.v_proj")
                L[key + "o"] = self._lin(sd, f"{lp}{att}.out_proj")
                L[key + "ln"] = ln(f"{lp}{att}_layer_norm")
            L["fc1"], L["fc2"], L["ln3"] = self._lin(sd, lp + "fc1"), self._lin(sd, lp + "fc2"), ln(lp + "final_layer_norm")
            self.layers.append(L)
        self.lm_head = sd["head.decoder.lm_head.weight"].float().to(self.dev).contiguous()

    # ---------------------------------------------------------------- op wrappers
    def _st(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream or None

    def _gemm(self, prec, A, lda, M, K, W, N, bias, act, res, ldr, out, ldc, c_off, out_step=None, out_step_stride=0):
        self.launches += 1
        _lib.check_op(self.lib.rdb_op_gemm(self.device, prec, A, lda, M, K, _lib.ptr(W), N, _lib.ptr(bias), act, res, ldr, out, ldc, c_off, self._st(),
                                           out_step, out_step_stride))

    def _esz(self):
        return 2 if self.prec == _lib.PREC_FP16 else 4

    def _conv(self, cv, x_ptr, n, h, w, ld, stride, pad, out_ptr, ldc, c_off, act, res=None, ldr=0, prec=None, same2=False):
        """dense conv (+folded BN, +act, +residual): x [n,h,w,cin] at x_ptr (pitch ld) -> out channel slice.  Returns (oh, ow)."""
        prec = self.prec if prec is None else prec
        k = cv.k
        if same2:                                   # 2x2 'same': zero pad right / bottom only
            oh, ow, pt = h, w, 0
        else:
            pt = pad
            oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        W = cv.w16 if prec == _lib.PREC_FP16 else cv.w
        M = n * oh * ow
        if k == 1 and stride == 1:
            self._gemm(prec, x_ptr, ld, M, cv.cin, W, cv.cout, cv.b, act, res, ldr, out_ptr, ldc, c_off)
            return oh, ow
        if prec == _lib.PREC_FP16 and self.implicit_conv and cv.cin % 8 == 0:
            # tcgen05 implicit GEMM: the taps are TMA boxes of the NHWC input (zero fill = padding), no im2col buffer
            self.launches += 1
            _lib.check_op(self.lib.rdb_op_conv_tc(self.device, x_ptr, n, h, w, cv.cin, ld, _lib.ptr(W), cv.cout, _lib.ptr(cv.b), act, k, k, stride, stride,
                                                  pt, pt, out_ptr, oh, ow, ldc, c_off, self._st()))
            assert res is None
            return oh, ow
        K = k * k * cv.cin
        col = self.torch.empty((M, K), dtype=self.torch.float16 if prec == _lib.PREC_FP16 else self.torch.float32, device=self.dev)
        self.launches += 1
        _lib.check_op(self.lib.rdb_op_im2col(self.device, prec, x_ptr, n, h, w, cv.cin, ld, k, k, stride, stride, pt, pt, oh, ow, col.data_ptr(), self._st()))
        self._gemm(prec, col.data_ptr(), K, M, K, W, cv.cout, cv.b, act, res, ldr, out_ptr, ldc, c_off)
        return oh, ow

    def _dw(self, cv, x_ptr, n, h, w, ld_in, stride, relu, out_ptr, ld_out, c_off):
        oh, ow = (h + 2 * ((cv.k - 1) // 2) - cv.k) // stride + 1, (w + 2 * ((cv.k - 1) // 2) - cv.k) // stride + 1
        self.launches += 1
        _lib.check_op(self.lib.rdb_op_dwconv(self.device, self.prec, x_ptr, n, h, w, cv.cout, ld_in, cv.k, stride, _lib.ptr(cv.w), _lib.ptr(cv.b), int(relu),
                                             out_ptr, oh, ow, ld_out, c_off, self._st()))
        return oh, ow

    # ---------------------------------------------------------------- encoder
    def encode(self, x):
        """x [B,1,H,W] float32 -> [B, (H/32)*(W/32), 2048] float32 device tensor (`last_hidden_state`)."""
        torch = self.torch
        with torch.cuda.device(self.dev):
            if isinstance(x, np.ndarray):
                x = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(self.dev)
            x = x.contiguous()
            B, c, H, W = x.shape
            assert c == 1 and H % 32 == 0 and W % 32 == 0
            es, adt, st = self._esz(), self.adt, self.stem
            # gray -> 3 channels (torch.repeat_interleave in the reference), NHWC fp32: stem1 always runs in fp32 (K = 27)
            x3 = torch.zeros((B * H * W, 4), dtype=torch.float32, device=self.dev)      # 4th channel = zero pad (K = 36, 16-byte rows)
            for ch in range(3):
                self.launches += 1
                _lib.check_op(self.lib.rdb_op_copy_cols(self.device, 0, 0, x.data_ptr(), B * H * W, 1, 1, x3.data_ptr(), 4, ch, self._st()))
            h1, w1 = H // 2, W // 2
            P1 = B * h1 * w1
            c1 = st["stem1"].cout
            e1f = torch.empty((P1, c1), dtype=torch.float32, device=self.dev)
            self._conv(st["stem1"], x3.data_ptr(), B, H, W, 4, 2, 1, e1f.data_ptr(), c1, 0, ACT_RELU, prec=_lib.PREC_FP32)
            if self.prec == _lib.PREC_FP16:
                e1 = torch.empty((P1, c1), dtype=adt, device=self.dev)
                self.launches += 1
                _lib.check_op(self.lib.rdb_op_copy_cols(self.device, 0, 1, e1f.data_ptr(), P1, c1, c1, e1.data_ptr(), c1, 0, self._st()))
            else:
                e1 = e1f
            a = torch.empty((P1, c1 // 2), dtype=adt, device=self.dev)
            self._conv(st["stem2a"], e1.data_ptr(), B, h1, w1, c1, 1, 0, a.data_ptr(), c1 // 2, 0, ACT_RELU, same2=True)
            cat = torch.empty((P1, 2 * c1), dtype=adt, device=self.dev)
            self._conv(st["stem2b"], a.data_ptr(), B, h1, w1, c1 // 2, 1, 0, cat.data_ptr(), 2 * c1, c1, ACT_RELU, same2=True)
            self.launches += 1
            _lib.check_op(self.lib.rdb_op_maxpool2x2s1(self.device, self.prec, e1.data_ptr(), B, h1, w1, c1, c1, cat.data_ptr(), 2 * c1, 0, self._st()))
            h2, w2 = h1 // 2, w1 // 2
            s3 = torch.empty((B * h2 * w2, c1), dtype=adt, device=self.dev)
            self._conv(st["stem3"], cat.data_ptr(), B, h1, w1, 2 * c1, 2, 1, s3.data_ptr(), c1, 0, ACT_RELU)
            # stem4 writes straight into the first block's wide buffer
            h, w = h2, w2
            cur = None                      # (tensor, ld, channels) of the current feature map
            pending = (st["stem4"], s3, c1)  # a 1x1 conv whose output location is decided by its consumer
            for si, stage in enumerate(self.stages):
                cin, mid, cout, blocks, down, light, k, layers = self.arch["stages"][si]
                if stage["down"] is not None:
                    # materialise the pending map, then depthwise 3x3 stride 2 (no activation)
                    t = torch.empty((B * h * w, cin), dtype=adt, device=self.dev)
                    self._emit(pending, B * h * w, t.data_ptr(), cin, 0)
                    nh, nw = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
                    total = cin + layers * mid
                    wide = torch.empty((B * nh * nw, total), dtype=adt, device=self.dev)
                    self._dw(stage["down"], t.data_ptr(), B, h, w, cin, 2, False, wide.data_ptr(), total, 0)
                    h, w = nh, nw
                    pending = None
                for bi, blk in enumerate(stage["blocks"]):
                    ic = cin if bi == 0 else cout
                    total = ic + layers * mid
                    P = B * h * w
                    if pending is not None:
                        wide = torch.empty((P, total), dtype=adt, device=self.dev)
                        self._emit(pending, P, wide.data_ptr(), total, 0)
                        pending = None
                    off, cprev = 0, ic
                    for (cv, dw) in blk["layers"]:
                        src = wide.data_ptr() + off * es
                        dst_off = off + cprev
                        if dw is None:
                            self._conv(cv, src, B, h, w, total, 1, (cv.k - 1) // 2, wide.data_ptr(), total, dst_off, ACT_RELU)
                        else:
                            t = torch.empty((P, mid), dtype=adt, device=self.dev)
                            self._conv(cv, src, B, h, w, total, 1, 0, t.data_ptr(), mid, 0, ACT_NONE)
                            self._dw(dw, t.data_ptr(), B, h, w, mid, 1, True, wide.data_ptr(), total, dst_off)
                        off, cprev = dst_off, mid
                    sq = torch.empty((P, cout // 2), dtype=adt, device=self.dev)
                    self._conv(blk["sq"], wide.data_ptr(), B, h, w, total, 1, 0, sq.data_ptr(), cout // 2, 0, ACT_RELU)
                    # the excitation conv (+ identity) lands in the NEXT consumer's buffer
                    pending = (blk["ex"], sq, cout // 2, (wide, total) if blk["identity"] else None)
            P = B * h * w
            out = torch.empty((P, self.arch["enc_dim"]), dtype=adt, device=self.dev)
            self._emit(pending, P, out.data_ptr(), self.arch["enc_dim"], 0)
            if self.prec == _lib.PREC_FP16:
                o32 = torch.empty((P, self.arch["enc_dim"]), dtype=torch.float32, device=self.dev)
                self.launches += 1
                _lib.check_op(self.lib.rdb_op_copy_cols(self.device, 1, 0, out.data_ptr(), P, self.arch["enc_dim"], self.arch["enc_dim"], o32.data_ptr(),
                                                        self.arch["enc_dim"], 0, self._st()))
                out = o32
            return out.view(B, h * w, self.arch["enc_dim"])

    def _emit(self, pending, P, out_ptr, ldc, c_off):
        """Run a deferred 1x1 conv (stem4 / aggregation_excitation_conv, ReLU, optional identity) into its consumer's buffer."""
        cv, src, cin = pending[0], pending[1], pending[2]
        ident = pending[3] if len(pending) > 3 else None
        res, ldr = (ident[0].data_ptr(), ident[1]) if ident is not None else (None, 0)
        W = cv.w16 if self.prec == _lib.PREC_FP16 else cv.w
        self._gemm(self.prec, src.data_ptr(), cin, P, cin, W, cv.cout, cv.b, ACT_RELU, res, ldr, out_ptr, ldc, c_off)

    # ---------------------------------------------------------------- decoder
    def _decoder_state(self, B, S):
        """Per (batch, encoder length) buffers + the captured CUDA graph of ONE decode step.  Every step-dependent quantity
        (position, cache row, attention length, token-table row) is read from a device counter, so the same graph replays
        for every step: ~100 kernel launches per token cost one graph launch."""
        key = (B, S)
        if key in self._dec:
            return self._dec[key]
        torch, a = self.torch, self.arch
        d, V = a["d_model"], a["vocab"]
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=self.dev)      # noqa: E731
        cap = self.max_new + 1
        st = dict(cap=cap, encp=new(B * S, d), cross=[(new(B * S, d), new(B * S, d)) for _ in self.layers],
                  kc=[new(B, cap, d) for _ in self.layers], vc=[new(B, cap, d) for _ in self.layers],
                  toks=torch.zeros((cap, B), dtype=torch.int64, device=self.dev), unfinished=torch.ones(B, dtype=torch.int32, device=self.dev),
                  has_eos=torch.zeros(B, dtype=torch.int32, device=self.dev), done=torch.zeros(cap, dtype=torch.int32, device=self.dev),
                  step=torch.zeros(1, dtype=torch.int32, device=self.dev), h=new(B, d), x=new(B, d), q=new(B, d), att=new(B, d), r1=new(B, d),
                  f=new(B, a["ffn"]), logits=new(B, V), arg=torch.empty(B, dtype=torch.int32, device=self.dev), val=new(B), graph=None)
        # the attention kernel's dynamic shared memory attribute must be set outside capture
        _lib.check_op(self.lib.rdb_op_attn_decode(self.device, st["q"].data_ptr(), st["kc"][0].data_ptr(), st["vc"][0].data_ptr(), B, 1, cap, a["heads"],
                                                  d // a["heads"], st["att"].data_ptr(), self._st(), st["step"].data_ptr()))
        self._decode_step(st, B, S)                # eager warm-up of every kernel on scratch state (no lazy init inside the capture)
        torch.cuda.synchronize(self.dev)
        if self.use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._decode_step(st, B, S)
            st["graph"] = g
        self._dec[key] = st
        return st

    def _decode_step(self, st, B, S):
        """One greedy step for the whole batch (MBartDecoderLayer x N, rec_unimernet_head.py:666-748; lm_head; argmax; bookkeeping)."""
        a, lib, dv, stm = self.arch, self.lib, self.device, self._st()
        d, H, V = a["d_model"], a["heads"], a["vocab"]
        hd, f32, cap = d // H, _lib.PREC_FP32, st["cap"]
        sp = st["step"].data_ptr()
        h, x, q, att, r1, f = st["h"], st["x"], st["q"], st["att"], st["r1"], st["f"]
        _lib.check_op(lib.rdb_op_embed(dv, st["toks"].data_ptr(), B, d, self.tok.data_ptr(), math.sqrt(d), self.pos.data_ptr(), 0, x.data_ptr(), stm, sp))
        _lib.check_op(lib.rdb_op_layernorm(dv, x.data_ptr(), B, d, self.ln_emb[0].data_ptr(), self.ln_emb[1].data_ptr(), 1e-5, h.data_ptr(), stm))
        for li, L in enumerate(self.layers):
            # self attention (pre-LN); k / v append to cache row `step`
            _lib.check_op(lib.rdb_op_layernorm(dv, h.data_ptr(), B, d, L["sln"][0].data_ptr(), L["sln"][1].data_ptr(), 1e-5, x.data_ptr(), stm))
            # q / k / v read the same LayerNorm output and write three different buffers: k and v go to side streams, which the
            # capture turns into parallel graph branches (three 13 us weight-streaming launches overlap instead of queueing)
            if self.qkv_parallel:
                main = self.torch.cuda.current_stream(self.dev)
                for s_ in self._side:
                    s_.wait_stream(main)
            self._gemm(f32, x.data_ptr(), d, B, d, L["sq"][0], d, L["sq"][1], ACT_NONE, None, 0, q.data_ptr(), d, 0)
            if self.qkv_parallel:
                with self.torch.cuda.stream(self._side[0]):
                    self._gemm(f32, x.data_ptr(), d, B, d, L["sk"][0], d, L["sk"][1], ACT_NONE, None, 0, st["kc"][li].data_ptr(), cap * d, 0, sp, d)
                with self.torch.cuda.stream(self._side[1]):
                    self._gemm(f32, x.data_ptr(), d, B, d, L["sv"][0], d, L["sv"][1], ACT_NONE, None, 0, st["vc"][li].data_ptr(), cap * d, 0, sp, d)
                for s_ in self._side:
                    main.wait_stream(s_)
            else:
                self._gemm(f32, x.data_ptr(), d, B, d, L["sk"][0], d, L["sk"][1], ACT_NONE, None, 0, st["kc"][li].data_ptr(), cap * d, 0, sp, d)
                self._gemm(f32, x.data_ptr(), d, B, d, L["sv"][0], d, L["sv"][1], ACT_NONE, None, 0, st["vc"][li].data_ptr(), cap * d, 0, sp, d)
            _lib.check_op(lib.rdb_op_attn_decode(dv, q.data_ptr(), st["kc"][li].data_ptr(), st["vc"][li].data_ptr(), B, 1, cap, H, hd, att.data_ptr(), stm, sp))
            self._gemm(f32, att.data_ptr(), d, B, d, L["so"][0], d, L["so"][1], ACT_NONE, h.data_ptr(), d, r1.data_ptr(), d, 0)
            # cross attention over the encoder tokens (K / V computed once per batch)
            _lib.check_op(lib.rdb_op_layernorm(dv, r1.data_ptr(), B, d, L["cln"][0].data_ptr(), L["cln"][1].data_ptr(), 1e-5, x.data_ptr(), stm))
            self._gemm(f32, x.data_ptr(), d, B, d, L["cq"][0], d, L["cq"][1], ACT_NONE, None, 0, q.data_ptr(), d, 0)
            _lib.check_op(lib.rdb_op_attn_decode(dv, q.data_ptr(), st["cross"][li][0].data_ptr(), st["cross"][li][1].data_ptr(), B, S, S, H, hd, att.data_ptr(), stm, None))
            self._gemm(f32, att.data_ptr(), d, B, d, L["co"][0], d, L["co"][1], ACT_NONE, r1.data_ptr(), d, h.data_ptr(), d, 0)
            # feed forward
            _lib.check_op(lib.rdb_op_layernorm(dv, h.data_ptr(), B, d, L["ln3"][0].data_ptr(), L["ln3"][1].data_ptr(), 1e-5, x.data_ptr(), stm))
            self._gemm(f32, x.data_ptr(), d, B, d, L["fc1"][0], a["ffn"], L["fc1"][1], ACT_GELU, None, 0, f.data_ptr(), a["ffn"], 0)
            self._gemm(f32, f.data_ptr(), a["ffn"], B, a["ffn"], L["fc2"][0], d, L["fc2"][1], ACT_NONE, h.data_ptr(), d, r1.data_ptr(), d, 0)
            h, r1 = r1, h
        _lib.check_op(lib.rdb_op_layernorm(dv, h.data_ptr(), B, d, self.ln_out[0].data_ptr(), self.ln_out[1].data_ptr(), 1e-5, x.data_ptr(), stm))
        self._gemm(f32, x.data_ptr(), d, B, d, self.lm_head, V, None, ACT_NONE, None, 0, st["logits"].data_ptr(), V, 0)
        _lib.check(lib.rdb_argmax_rows(dv, st["logits"].data_ptr(), B, V, st["arg"].data_ptr(), st["val"].data_ptr(), stm))
        _lib.check_op(lib.rdb_op_greedy_step(dv, st["arg"].data_ptr(), B, 0, a["eos"], a["pad"], st["toks"].data_ptr(), st["unfinished"].data_ptr(),
                                             st["has_eos"].data_ptr(), st["done"].data_ptr(), stm, sp, a["forced_eos_len"]))

    def generate(self, enc):
        """enc [B,S,enc_dim] float32 device tensor -> ids [B, L] int64 numpy (start token first), as generate_export returns."""
        torch, a = self.torch, self.arch
        with torch.cuda.device(self.dev):
            B, S, E = enc.shape
            assert B <= 32, "decode batches are at most 32 rows (split larger batches)"
            d = a["d_model"]
            f32 = _lib.PREC_FP32
            st = self._decoder_state(B, S)
            self._gemm(f32, enc.data_ptr(), E, B * S, E, self.proj[0], d, self.proj[1], ACT_NONE, None, 0, st["encp"].data_ptr(), d, 0)
            for L, (ck, cv_) in zip(self.layers, st["cross"]):
                self._gemm(f32, st["encp"].data_ptr(), d, B * S, d, L["ck"][0], d, L["ck"][1], ACT_NONE, None, 0, ck.data_ptr(), d, 0)
                self._gemm(f32, st["encp"].data_ptr(), d, B * S, d, L["cv"][0], d, L["cv"][1], ACT_NONE, None, 0, cv_.data_ptr(), d, 0)
            st["toks"].zero_()
            st["toks"][0] = a["start"]
            st["unfinished"].fill_(1)
            st["has_eos"].zero_()
            st["done"].zero_()
            st["step"].zero_()
            steps = 0
            per_step = 15 * len(self.layers) + 5
            while steps < self.max_new:
                if st["graph"] is not None:
                    st["graph"].replay()
                else:
                    self._decode_step(st, B, S)
                self.launches += per_step
                steps += 1
                if steps % self.sync_every == 0 and bool(st["done"][steps].item()):
                    break
            dn = st["done"][:steps + 1].cpu().numpy()
            first = np.nonzero(dn)[0]
            L_out = int(first[0]) if len(first) else steps          # generate_export stops right after the step that completed every row
            return st["toks"][:L_out + 1].t().contiguous().cpu().numpy()

    def __call__(self, x):
        """x [B,1,H,W] -> ids; batches larger than 32 are decoded 32 rows at a time (ragged tails padded by the caller's batching)."""
        n = x.shape[0]
        if n <= 32:
            return self.generate(self.encode(x))
        outs = [self.generate(self.encode(x[i:i + 32])) for i in range(0, n, 32)]
        L = max(o.shape[1] for o in outs)
        return np.concatenate([np.pad(o, ((0, 0), (0, L - o.shape[1])), constant_values=self.arch["pad"]) for o in outs], 0)


class B200FormulaSession:
    """The reference's formula `InferSession` protocol (rapid_formula_self/inference_engine/torch.py:25-131):
    `__call__(np [B,1,384,384] f32) -> [ids [B,L]]`."""

    def __init__(self, engine):
        self.engine = engine

    def __call__(self, img):
        return [self.engine(np.asarray(img, np.float32))]


class B200FormulaModel:
    """F1 + the CustomBaseModel plugin surface: `batch_predict(images RGB uint8, batch_size) -> list[str]`
    (rapid_formula_model.py:34-41; plugin contract rapid_doc/model/custom/__init__.py:4-20, call site
    backend/pipeline/batch_analyze.py:272-284).  `decode(ids) -> str` is the tokenizer + LaTeX normalisation (F5, UniMERNetDecode);
    the tokenizer file ships with the checkpoint, so it is injected — without it the token ids are returned as a space-joined
    string (explicitly marked), never silently."""

    def __init__(self, engine, decode=None, batch_size=32):
        self.engine, self.decode, self.batch_size = engine, decode, batch_size
        self.pre = FormulaPreProcess(engine.arch["input_size"])

    def predict_ids(self, images, batch_size=None):
        bs = int(batch_size or self.batch_size)
        out = []
        for b0 in range(0, len(images), bs):
            x = np.concatenate(self.pre(images[b0:b0 + bs]), axis=0)
            ids = self.engine(x)
            out.extend(list(ids))
        return out

    def batch_predict(self, images, batch_size=None, **kwargs):
        ids = self.predict_ids(images, batch_size)
        if self.decode is None:
            return ["<ids> " + " ".join(str(int(t)) for t in row) for row in ids]
        return [self.decode(row) for row in ids]
