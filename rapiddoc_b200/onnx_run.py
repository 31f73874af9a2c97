"""A B200 executor for the small CNN graphs RapidDoc ships as ONNX files and runs through onnxruntime on the CPU:

  rapid_orientation.onnx          PP-LCNet x1.0, 4-way table-orientation classifier  — `OrtInferSession.__call__`
                                  (rapid_doc/model/orientation/rapid_orientation/utils.py:47-52, called rapid_orientation.py:48)
  pp-ocrv4_mobile_seal_det.onnx   PP-LCNetV3 + RSE-FPN + DB head of the seal-text detector

The graph is read by `onnx_lite` (no onnx / onnxruntime package), compiled once into a launch plan over the C-ABI ops of
include/rapiddoc_b200.h (NHWC fp32: rdb_op_gemm / rdb_op_im2col / rdb_op_dwconv for the convolutions, rdb_op_chain for every
run of BatchNormalization / bias / learnable-affine / activation nodes between two convolutions — ONE launch per run instead
of one per node — rdb_op_global_avgpool + rdb_op_mul_gate for squeeze-excite, rdb_op_resize_nearest, rdb_op_depth_to_space for
ConvTranspose, rdb_op_softmax_rows) and executed on the caller's CUDA stream.  There is no CPU path: without the CUDA library
`_lib.load()` raises.  tests/test_onnx_run.py compares it with `oracle/onnx_ref.py` (torch CPU), which is itself pinned to
OpenCV's DNN importer on the same files.
"""
import ctypes

import numpy as np

from . import _lib, onnx_lite

CH_AFFINE, CH_AFFINE_VEC, CH_RELU, CH_HSIG, CH_HSWISH, CH_SIGMOID = range(6)
MAX_STEPS = 8
ACT_NONE = 0


class _T:
    """A device activation: fp32 NHWC rows (rows = n*h*w) of `c` channels starting at column `off` of buf [rows, ld] — a plain
    buffer (off 0, ld == c) or a channel slice of a wider one (a Concat output its producer wrote in place) — plus a pending chain
    of per-element steps not yet applied."""
    __slots__ = ("buf", "n", "h", "w", "c", "steps", "off")

    def __init__(self, buf, n, h, w, c, steps=(), off=0):
        self.buf, self.n, self.h, self.w, self.c, self.steps, self.off = buf, n, h, w, c, tuple(steps), off

    @property
    def rows(self):
        return self.n * self.h * self.w

    @property
    def ptr(self):
        return self.buf.data_ptr() + 4 * self.off

    @property
    def ld(self):
        return self.buf.shape[1]

    @property
    def dense(self):
        return self.off == 0 and self.buf.shape[1] == self.c


class OnnxCnn:
    def __init__(self, path, device=0, compile_only=False, precision="fp32"):
        """compile_only: read the file and build the launch plan without touching a device (plan inspection in CPU tests).
        precision: "fp32" = fp32 SIMT GEMMs (exact mode); "tf32" = the same fp32 buffers multiplied on the tensor cores
        (tcgen05 kind::tf32, RDB_PREC_TF32) for every GEMM with more than 32 rows."""
        assert precision in ("fp32", "tf32")
        self.precision = precision
        self.gemm_prec = _lib.PREC_TF32 if precision == "tf32" else _lib.PREC_FP32
        self.direct_stem = True
        import torch
        self.torch, self.device, self.lib = torch, int(device), _lib.load()
        if not compile_only:
            if self.lib.rdb_device_count() <= self.device:
                raise _lib.B200Error(f"no sm_100 CUDA device {self.device} for the ONNX executor (there is no CPU fallback)")
            self.dev = torch.device("cuda", self.device)
        g = onnx_lite.load(path)
        self.graph, self.path = g, path
        self.meta = getattr(g, "meta", {})
        self.launches = 0
        self._const = {}          # name -> numpy constant (initializers + folded shape arithmetic)
        self._dev = {}            # cache key -> device tensor (weights / per-channel vectors)
        self._alias, self._plans, self._folded = {}, {}, {}
        self.consts = dict(g.init)                    # initializers + Constant nodes (Paddle2ONNX exports the weights either way)
        for n in g.nodes:
            if n.op == "Constant":
                self.consts[n.outputs[0]] = np.asarray(n.attrs["value"])
        self.nodes = self._rewrite(g)
        self._place, self._cat = self._plan_concats(self.nodes), {}
        self.input_name, self.output_name = g.inputs[0], g.outputs[0]

    # ---------------------------------------------------------------- compile
    def _rewrite(self, g):
        """Drop Identity nodes; fuse Mul(HardSigmoid(y), y) into one HardSwish-with-(alpha, beta) node."""
        alias = {}
        nodes = []
        for n in g.nodes:
            if n.op == "Identity":
                alias[n.outputs[0]] = alias.get(n.inputs[0], n.inputs[0])
            elif n.op == "Constant":
                continue
            else:
                nodes.append(onnx_lite.Node(n.op, [alias.get(i, i) for i in n.inputs], list(n.outputs), dict(n.attrs), n.name))
        self._alias = alias
        uses = {}
        for n in nodes:
            for i in n.inputs:
                uses[i] = uses.get(i, 0) + 1
        prod = {o: n for n in nodes for o in n.outputs}
        dead = set()
        for n in nodes:
            if n.op != "Mul" or len(n.inputs) != 2:
                continue
            for a, b in (n.inputs, n.inputs[::-1]):
                p = prod.get(a)
                if p is not None and p.op == "HardSigmoid" and p.inputs[0] == b and uses.get(a, 0) == 1:
                    n.op, n.inputs = "HardSwishAB", [b]
                    n.attrs = {"alpha": p.attrs.get("alpha", 0.2), "beta": p.attrs.get("beta", 0.5)}
                    dead.add(id(p))
                    break
        nodes = [n for n in nodes if id(n) not in dead]
        return self._fuse_conv_epilogues(nodes)

    def _fuse_conv_epilogues(self, nodes):
        """Conv -> BatchNormalization -> (HardSwish | Relu), and Conv -> Add(per-channel constant) -> Relu, collapse into the conv
        launch: BN scale folded into the packed weights, shift / bias into the GEMM bias, the activation into its epilogue."""
        uses = {}
        for n in nodes:
            for i in n.inputs:
                uses[i] = uses.get(i, 0) + 1
        for o in self.graph.outputs:
            uses[o] = uses.get(o, 0) + 1
        consumer = {}
        for n in nodes:
            for i in n.inputs:
                consumer.setdefault(i, n)
        dead = set()
        for n in nodes:
            if n.op != "Conv":
                continue
            def only_next(ops):
                out = n.outputs[0]
                m = consumer.get(out)
                return m if m is not None and uses.get(out, 0) == 1 and m.op in ops and m.inputs[0] == out and id(m) not in dead else None
            m = only_next(("BatchNormalization",))
            if m is not None:
                n.attrs["fold_bn"] = (list(m.inputs[1:5]), float(m.attrs.get("epsilon", 1e-5)))
                n.outputs = list(m.outputs)
                dead.add(id(m))
            else:
                m = only_next(("Add",))
                if m is not None and len(n.inputs) < 3:
                    c = self.consts.get(m.inputs[1])
                    co = self.consts[n.inputs[1]].shape[0]
                    if c is not None and np.asarray(c).size == co:
                        n.attrs["fold_bias"] = m.inputs[1]
                        n.outputs = list(m.outputs)
                        dead.add(id(m))
            m = only_next(("HardSwish", "Relu"))
            if m is not None:
                n.attrs["fold_act"] = 6 if m.op == "HardSwish" else 1
                n.outputs = list(m.outputs)
                dead.add(id(m))
        return [n for n in nodes if id(n) not in dead]

    def _weight(self, key, make):
        t = self._dev.get(key)
        if t is None:
            t = self.torch.from_numpy(np.ascontiguousarray(make(), np.float32)).to(self.dev)
            self._dev[key] = t
        return t

    # ---------------------------------------------------------------- launch helpers
    def _st(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream or None

    def _new(self, rows, c):
        return self.torch.empty((rows, c), dtype=self.torch.float32, device=self.dev)

    def _chain(self, t, out=None, ld_out=None, c_off=0):
        """Apply t's pending steps (possibly none: a strided copy) into `out` (a new [rows, c] buffer by default)."""
        if out is None:
            out, ld_out = self._new(t.rows, t.c), t.c
        n = len(t.steps)
        kinds = (ctypes.c_int32 * max(n, 1))(*[s[0] for s in t.steps])
        a = (ctypes.c_float * max(n, 1))(*[s[1] for s in t.steps])
        b = (ctypes.c_float * max(n, 1))(*[s[2] for s in t.steps])
        va = (ctypes.c_void_p * max(n, 1))(*[s[3].data_ptr() if s[3] is not None else None for s in t.steps])
        vb = (ctypes.c_void_p * max(n, 1))(*[s[4].data_ptr() if s[4] is not None else None for s in t.steps])
        self.launches += 1
        _lib.check_op(self.lib.rdb_op_chain(self.device, t.ptr, t.rows, t.c, t.ld, kinds, a, b, va, vb, n, out.data_ptr(), ld_out, c_off,
                                            self._st()))
        return out

    def _mat(self, t, dense=False):
        """The activation with its pending steps applied (materialised); dense: also as a plain [rows, c] buffer."""
        if not t.steps and (t.dense or not dense):
            return t
        return _T(self._chain(t), t.n, t.h, t.w, t.c)

    def _push(self, t, step):
        if len(t.steps) >= MAX_STEPS:
            t = self._mat(t)
        return _T(t.buf, t.n, t.h, t.w, t.c, t.steps + (step,), t.off)

    def _out(self, name, rows, c):
        """Where a conv / resize writes: its own [rows, c] buffer, or — when the compile pass placed it — its channel slice of
        the Concat output it feeds (allocated by the first producer).  -> (buffer, row pitch, column offset)"""
        place = self._place.get(name)
        if place is None:
            return self._new(rows, c), c, 0
        cat, off, ctot = place
        buf = self._cat.get(cat)
        if buf is None:
            buf = self._cat[cat] = self._new(rows, ctot)
        assert buf.shape == (rows, ctot)
        return buf, ctot, off

    def _plan_concats(self, nodes):
        """Channel-count inference + placement: a Conv / Resize output that feeds a channel Concat is written straight into its
        slice of the Concat buffer (every op reads pitched slices), so the Concat launches no copy for it."""
        chan, prod = {self.graph.inputs[0]: 4}, {}
        for n in nodes:
            for o in n.outputs:
                prod[o] = n
            c = None
            if n.op == "Conv":
                c = int(self.consts[n.inputs[1]].shape[0])
            elif n.op == "ConvTranspose":
                c = int(self.consts[n.inputs[1]].shape[1])
            elif n.op == "MatMul" and n.inputs[1] in self.consts:
                c = int(self.consts[n.inputs[1]].shape[1])
            elif n.op == "Concat" and n.attrs.get("axis") == 1:
                cs = [chan.get(i) for i in n.inputs]
                c = sum(cs) if all(v is not None for v in cs) else None
            else:
                c = next((chan[i] for i in n.inputs if i in chan and chan[i] is not None), None)
            for o in n.outputs:
                chan[o] = c
        place = {}
        for n in nodes:
            if n.op != "Concat" or n.attrs.get("axis") != 1:
                continue
            cs = [chan.get(i) for i in n.inputs]
            if any(v is None or v % 4 for v in cs):
                continue
            off = 0
            for i, c in zip(n.inputs, cs):
                p = prod.get(i)
                if p is not None and p.op in ("Conv", "Resize") and i not in place:
                    place[i] = (n.outputs[0], off, sum(cs))
                off += c
        return place

    def _gemm(self, A, lda, M, K, W, N, bias, out, ldc, c_off=0, act=ACT_NONE):
        self.launches += 1
        _lib.check_op(self.lib.rdb_op_gemm(self.device, self.gemm_prec, A, lda, M, K, W.data_ptr(), N, bias.data_ptr() if bias is not None else None, act,
                                           None, 0, out, ldc, c_off, self._st(), None, 0))

    # ---------------------------------------------------------------- ops
    def _conv(self, node, x):
        wname = node.inputs[1]
        W = self.consts[wname]
        co, cig, kh, kw = W.shape
        group = node.attrs.get("group", 1)
        sh, sw = node.attrs.get("strides", [1, 1])
        pads = node.attrs.get("pads", [0, 0, 0, 0])
        assert kh == kw and sh == sw and len(set(pads)) == 1 and set(node.attrs.get("dilations", [1, 1])) == {1}, f"unsupported conv geometry {node}"
        p = pads[0]
        act = int(node.attrs.get("fold_act", ACT_NONE))
        scale = None
        if "fold_bn" in node.attrs:
            names, eps = node.attrs["fold_bn"]
            if ("bnb", wname) not in self._dev or wname not in self._folded:
                gamma, beta, mean, var = (self.consts[i].astype(np.float64) for i in names)
                scale = gamma / np.sqrt(var + eps)
                shift = beta - mean * scale
                if len(node.inputs) > 2 and node.inputs[2]:
                    shift = shift + self.consts[node.inputs[2]].astype(np.float64) * scale
            else:
                shift = None
            bias = self._weight(("bnb", wname), lambda: shift)
            if wname not in self._folded:
                self._folded[wname] = (W.astype(np.float64) * scale.reshape(-1, 1, 1, 1)).astype(np.float32)
            W = self._folded[wname]
        elif "fold_bias" in node.attrs:
            bias = self._weight(("fb", node.attrs["fold_bias"]), lambda: np.asarray(self.consts[node.attrs["fold_bias"]]).reshape(-1))
        else:
            bias = self._weight(("b", node.inputs[2]), lambda: self.consts[node.inputs[2]]) if len(node.inputs) > 2 and node.inputs[2] else None
        x = self._mat(x)
        oh, ow = (x.h + 2 * p - kh) // sh + 1, (x.w + 2 * p - kw) // sw + 1
        if group == 1:
            if x.c % 4 or x.ld % 4 or x.off % 4:  # the GEMM reads 16-byte vectors: pad the channels to a multiple of 4
                cp = (x.c + 3) // 4 * 4
                padded = self.torch.zeros((x.rows, cp), dtype=self.torch.float32, device=self.dev)
                self._chain(x, padded, cp, 0)
                x = _T(padded, x.n, x.h, x.w, cp)
            cin = x.c                                       # (the network input is stored padded the same way)
            assert cig <= cin and cin % 4 == 0 and x.ld % 4 == 0

            def pack():
                w = np.zeros((co, kh, kw, cin), np.float32)
                w[..., :cig] = W.transpose(0, 2, 3, 1)
                return w.reshape(co, kh * kw * cin)
            Wd = self._weight(("w", wname, cin), pack)
            M = x.n * oh * ow
            out, ldc, off = self._out(node.outputs[0], M, co)
            if kh == 1 and sh == 1 and p == 0:
                self._gemm(x.ptr, x.ld, M, cin, Wd, co, bias, out.data_ptr(), ldc, off, act=act)
            elif kh == 3 and cin == 4 and x.ld == 4 and co == 16 and self.direct_stem:
                self.launches += 1              # the network's first layer: direct conv, no 9x im2col buffer (bit-identical)
                _lib.check_op(self.lib.rdb_op_conv3x3_c4(self.device, x.ptr, x.n, x.h, x.w, Wd.data_ptr(), co, bias.data_ptr() if bias is not None else None, act,
                                                         sh, p, out.data_ptr(), oh, ow, ldc, off, self._st()))
            else:
                K = kh * kw * cin
                col = self._new(M, K)
                self.launches += 1
                _lib.check_op(self.lib.rdb_op_im2col(self.device, _lib.PREC_FP32, x.ptr, x.n, x.h, x.w, cin, x.ld, kh, kw, sh, sw, p, p, oh, ow,
                                                     col.data_ptr(), self._st()))
                self._gemm(col.data_ptr(), K, M, K, Wd, co, bias, out.data_ptr(), ldc, off, act=act)
            return _T(out, x.n, oh, ow, co, off=off)
        assert group == x.c == co and cig == 1 and p == (kh - 1) // 2 and (kh & 1), f"only depthwise grouped convs are supported: {node}"
        Wd = self._weight(("dw", wname), lambda: W.reshape(co, kh * kw).T)
        if bias is None:
            bias = self._weight(("zeros", co), lambda: np.zeros(co, np.float32))
        out, ldc, off = self._out(node.outputs[0], x.n * oh * ow, co)
        self.launches += 1
        _lib.check_op(self.lib.rdb_op_dwconv(self.device, _lib.PREC_FP32, x.ptr, x.n, x.h, x.w, co, x.ld, kh, sh, Wd.data_ptr(), bias.data_ptr(),
                                             act, out.data_ptr(), oh, ow, ldc, off, self._st()))
        return _T(out, x.n, oh, ow, co, off=off)

    def _conv_transpose(self, node, x):
        W = self._const_of(node.inputs[1])
        ci, co, kh, kw = W.shape
        s = node.attrs.get("strides", [1, 1])
        assert kh == kw == s[0] == s[1] and set(node.attrs.get("pads", [0, 0, 0, 0])) == {0} and node.attrs.get("group", 1) == 1, f"unsupported ConvTranspose {node}"
        k = kh
        x = self._mat(x)
        assert ci == x.c and ci % 4 == 0
        Wd = self._weight(("wt", node.inputs[1]), lambda: W.transpose(2, 3, 1, 0).reshape(k * k * co, ci))
        bias = None
        if len(node.inputs) > 2 and node.inputs[2]:
            bias = self._weight(("bt", node.inputs[2]), lambda: np.tile(self._const_of(node.inputs[2]).reshape(-1), k * k))
        M = x.rows
        tmp = self._new(M, k * k * co)
        self._gemm(x.ptr, x.ld, M, ci, Wd, k * k * co, bias, tmp.data_ptr(), k * k * co)
        out = self._new(M * k * k, co)
        self.launches += 1
        _lib.check_op(self.lib.rdb_op_depth_to_space(self.device, tmp.data_ptr(), x.n, x.h, x.w, co, k, out.data_ptr(), self._st()))
        return _T(out, x.n, x.h * k, x.w * k, co)

    def _const_of(self, name):
        if name in self.consts:
            return self.consts[name]
        return self._const.get(name)

    def _affine_const(self, t, c, mul):
        """x * c or x + c for a constant c: scalar, or per channel ([C], [1,C,1,1])."""
        c = np.asarray(c, np.float32)
        if c.size == 1:
            v = float(c.reshape(-1)[0])
            return self._push(t, (CH_AFFINE, v, 0.0, None, None) if mul else (CH_AFFINE, 1.0, v, None, None))
        assert c.size == t.c and (c.ndim == 1 or (c.ndim == 4 and c.shape[1] == t.c)), f"unsupported broadcast {c.shape} on C={t.c}"
        vec = self._weight(("vec", c.tobytes()), lambda: c.reshape(-1))
        return self._push(t, (CH_AFFINE_VEC, 0.0, 0.0, vec, None) if mul else (CH_AFFINE_VEC, 0.0, 0.0, None, vec))

    def _binary(self, node, env, mul):
        a, b = node.inputs
        ca, cb = self._const_of(a), self._const_of(b)
        if ca is not None and cb is not None:
            self._const[node.outputs[0]] = ca * cb if mul else ca + cb
            return None
        if ca is not None or cb is not None:
            t, c = (env[b], ca) if ca is not None else (env[a], cb)
            return self._affine_const(t, c, mul)
        ta, tb = env[a], env[b]
        if mul:
            gate, x = (ta, tb) if ta.h * ta.w == 1 and tb.h * tb.w > 1 else (tb, ta)
            assert gate.h * gate.w == 1 and gate.c == x.c and gate.n == x.n, f"unsupported Mul of two activations: {node}"
            gate, x = self._mat(gate, dense=True), self._mat(x)
            assert gate.ld == gate.c
            out = self._new(x.rows, x.c)
            self.launches += 1
            _lib.check_op(self.lib.rdb_op_mul_gate(self.device, x.ptr, gate.ptr, x.n, x.h * x.w, x.c, x.ld, out.data_ptr(), x.c,
                                                   self._st()))
            return _T(out, x.n, x.h, x.w, x.c)
        ta, tb = self._mat(ta, dense=True), self._mat(tb, dense=True)
        assert (ta.n, ta.h, ta.w, ta.c) == (tb.n, tb.h, tb.w, tb.c) and ta.ld == ta.c and tb.ld == tb.c, f"Add of unequal shapes: {node}"
        out = self._new(ta.rows, ta.c)
        self.launches += 1
        _lib.check_op(self.lib.rdb_op_add(self.device, ta.ptr, tb.ptr, out.data_ptr(), ta.rows * ta.c, self._st()))
        return _T(out, ta.n, ta.h, ta.w, ta.c)

    def _bn(self, node, x):
        gamma, beta, mean, var = (self.consts[i].astype(np.float64) for i in node.inputs[1:5])
        eps = float(node.attrs.get("epsilon", 1e-5))
        key = ("bn", node.inputs[1])
        scale = self._weight(key + ("s",), lambda: gamma / np.sqrt(var + eps))
        shift = self._weight(key + ("t",), lambda: beta - mean * gamma / np.sqrt(var + eps))
        return self._push(x, (CH_AFFINE_VEC, 0.0, 0.0, scale, shift))

    # ---------------------------------------------------------------- run
    def _closure(self, targets):
        """The nodes (in graph order) the named tensors depend on."""
        prod = {o: n for n in self.nodes for o in n.outputs}
        need, stack = set(), [self._alias.get(t, t) for t in targets]
        while stack:
            t = stack.pop()
            n = prod.get(t)
            if n is None or id(n) in need:
                continue
            need.add(id(n))
            stack.extend(i for i in n.inputs if i)
        return [n for n in self.nodes if id(n) in need]

    def features(self, x, name, nhwc4=None):
        """Run only what the tensor `name` needs and return it on the device as (buffer [n*h*w, C] NHWC fp32, n, h, w, C).
        nhwc4=(n, h, w): `x` is already the network input on the device as [n*h*w, 4] fp32 NHWC (3 channels + a zero one)."""
        torch = self.torch
        with torch.cuda.device(self.dev):
            if nhwc4 is not None:
                n, h, w = nhwc4
                self._const, self._flat_out, self._cat = {}, False, {}
                env = {self.input_name: _T(x.view(n * h * w, 4), n, h, w, 4)}
            else:
                env = self._feed(x)
            plan = self._plans.get(name)
            if plan is None:
                plan = self._plans[name] = self._closure([name])
            for node in plan:
                out = self._run_node(node, env)
                if out is not None:
                    env[node.outputs[0]] = out
            t = self._mat(env[self._alias.get(name, name)], dense=True)
            assert t.ld == t.c
            return t.buf, t.n, t.h, t.w, t.c

    def _feed(self, x):
        torch = self.torch
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(self.dev, non_blocking=False)
        n, c, h, w = x.shape
        cp = (c + 3) // 4 * 4
        xin = torch.zeros((n, h, w, cp), dtype=torch.float32, device=self.dev)
        xin[..., :c] = x.permute(0, 2, 3, 1)
        self._const, self._flat_out, self._cat = {}, False, {}
        return {self.input_name: _T(xin.view(n * h * w, cp), n, h, w, cp)}

    def __call__(self, x):
        """x [n, 3, h, w] float32 (numpy or a CUDA tensor) -> numpy output of the graph (NCHW / [n, classes])."""
        torch = self.torch
        with torch.cuda.device(self.dev):
            env = self._feed(x)
            for node in self.nodes:
                out = self._run_node(node, env)
                if out is not None:
                    env[node.outputs[0]] = out
            name = self._alias.get(self.output_name, self.output_name)
            t = self._mat(env[name], dense=True)
            y = t.buf.view(t.n, t.h, t.w, t.ld)[..., :t.c]
            y = y.permute(0, 3, 1, 2).contiguous().cpu().numpy()
            return y.reshape(t.n, t.c) if self._flat_out and t.h * t.w == 1 else y

    def _host_fold(self, node, env):
        """Shape arithmetic on the host: nodes whose inputs are all constants (or `Shape` of an activation)."""
        op, a = node.op, node.attrs
        if op == "Shape":
            t = env.get(node.inputs[0])
            if t is None:
                return False
            self._const[node.outputs[0]] = np.array([t.n, t.c, t.h, t.w], np.int64)
            return True
        if op not in ("Cast", "Slice", "Squeeze", "Unsqueeze", "Reshape", "Concat", "Gather"):
            return False
        vals = [self._const_of(i) if i else None for i in node.inputs]
        if any(v is None and i for v, i in zip(vals, node.inputs)):
            return False
        v = [np.asarray(x) if x is not None else None for x in vals]
        if op == "Cast":
            y = v[0].astype({1: np.float32, 6: np.int32, 7: np.int64, 9: np.bool_, 11: np.float64}[a["to"]])
        elif op == "Slice":
            axes = v[3].reshape(-1) if len(v) > 3 and v[3] is not None else range(v[1].size)
            y = v[0]
            for st, en, ax in zip(v[1].reshape(-1), v[2].reshape(-1), axes):
                y = np.take(y, range(*slice(int(st), int(en)).indices(y.shape[int(ax)])), axis=int(ax))
        elif op == "Squeeze":
            y = np.squeeze(v[0], axis=tuple(int(x) for x in v[1].reshape(-1)) if len(v) > 1 and v[1] is not None else None)
        elif op == "Unsqueeze":
            y = v[0]
            for ax in sorted(int(x) for x in v[1].reshape(-1)):
                y = np.expand_dims(y, ax)
        elif op == "Reshape":
            y = v[0].reshape([int(x) for x in v[1].reshape(-1)])
        elif op == "Gather":
            y = np.take(v[0], v[1].astype(np.int64), axis=a.get("axis", 0))
        else:
            y = np.concatenate([np.atleast_1d(x) for x in v], axis=a.get("axis", 0))
        self._const[node.outputs[0]] = y
        return True

    def _run_node(self, node, env):
        op = node.op
        if self._host_fold(node, env):
            return None
        if op == "Conv":
            return self._conv(node, env[node.inputs[0]])
        if op == "ConvTranspose":
            return self._conv_transpose(node, env[node.inputs[0]])
        if op == "BatchNormalization":
            return self._bn(node, env[node.inputs[0]])
        if op == "Relu":
            return self._push(env[node.inputs[0]], (CH_RELU, 0.0, 0.0, None, None))
        if op == "Sigmoid":
            return self._push(env[node.inputs[0]], (CH_SIGMOID, 0.0, 0.0, None, None))
        if op == "HardSigmoid":
            return self._push(env[node.inputs[0]], (CH_HSIG, float(node.attrs.get("alpha", 0.2)), float(node.attrs.get("beta", 0.5)), None, None))
        if op == "HardSwish":
            return self._push(env[node.inputs[0]], (CH_HSWISH, 1.0 / 6.0, 0.5, None, None))
        if op == "HardSwishAB":
            return self._push(env[node.inputs[0]], (CH_HSWISH, float(node.attrs["alpha"]), float(node.attrs["beta"]), None, None))
        if op in ("Add", "Mul"):
            return self._binary(node, env, op == "Mul")
        if op == "GlobalAveragePool":
            x = self._mat(env[node.inputs[0]])
            out = self._new(x.n, x.c)
            self.launches += 1
            _lib.check_op(self.lib.rdb_op_global_avgpool(self.device, x.ptr, x.n, x.h * x.w, x.c, x.ld, out.data_ptr(), self._st()))
            return _T(out, x.n, 1, 1, x.c)
        if op == "Resize":
            assert node.attrs.get("mode") == "nearest" and node.attrs.get("coordinate_transformation_mode") == "asymmetric" \
                and node.attrs.get("nearest_mode", "floor") == "floor", f"unsupported Resize {node.attrs}"
            x = self._mat(env[node.inputs[0]])
            sizes = self._const_of(node.inputs[3]) if len(node.inputs) > 3 and node.inputs[3] else None
            if sizes is not None and np.asarray(sizes).size == 4:
                oh, ow = (int(v) for v in np.asarray(sizes).reshape(-1)[2:])
            else:
                scales = np.asarray(self._const_of(node.inputs[2]), np.float32)
                assert scales[0] == scales[1] == 1
                oh, ow = int(x.h * scales[2]), int(x.w * scales[3])
            out, ldc, off = self._out(node.outputs[0], x.n * oh * ow, x.c)
            self.launches += 1
            _lib.check_op(self.lib.rdb_op_resize_nearest(self.device, x.ptr, x.n, x.h, x.w, x.c, x.ld, oh, ow, out.data_ptr(), ldc, off, self._st()))
            return _T(out, x.n, oh, ow, x.c, off=off)
        if op == "Concat":
            if all(self._const_of(i) is not None for i in node.inputs):
                self._const[node.outputs[0]] = np.concatenate([np.atleast_1d(self._const_of(i)) for i in node.inputs], axis=node.attrs.get("axis", 0))
                return None
            assert node.attrs.get("axis") == 1, "only channel Concat is supported"
            parts = [env[i] for i in node.inputs]
            ctot = sum(p.c for p in parts)
            p0 = parts[0]
            out = self._cat.get(node.outputs[0])             # already holds the slices its producers wrote in place
            if out is None:
                out = self._new(p0.rows, ctot)
            assert out.shape == (p0.rows, ctot)
            off = 0
            for p in parts:
                assert (p.n, p.h, p.w) == (p0.n, p0.h, p0.w)
                if not (p.buf is out and p.off == off and not p.steps):
                    self._chain(p, out, ctot, off)          # pending steps (or a plain copy) straight into the channel slice
                off += p.c
            return _T(out, p0.n, p0.h, p0.w, ctot)
        if op == "Shape":
            t = env[node.inputs[0]]
            self._const[node.outputs[0]] = np.array([t.n, t.c, t.h, t.w], np.int64)
            return None
        if op == "Slice":
            data, starts, ends = (self._const_of(i) for i in node.inputs[:3])
            axes = self._const_of(node.inputs[3]) if len(node.inputs) > 3 else np.array([0])
            assert data is not None and int(np.asarray(axes).reshape(-1)[0]) == 0, "Slice is only supported on shape vectors"
            self._const[node.outputs[0]] = data[int(starts.reshape(-1)[0]): int(ends.reshape(-1)[0])]
            return None
        if op == "Reshape":
            t = env[node.inputs[0]]
            shape = self._const_of(node.inputs[1])
            assert t.h * t.w == 1 and len(shape) == 2, "Reshape is only supported as the [n,C,1,1] -> [n,C] flatten"
            self._flat_out = True
            return t
        if op == "MatMul":
            t = self._mat(env[node.inputs[0]])
            W = self._const_of(node.inputs[1])
            assert t.h * t.w == 1 and W.shape[0] == t.c and t.c % 4 == 0
            Wd = self._weight(("mm", node.inputs[1]), lambda: W.T)
            out = self._new(t.n, W.shape[1])
            self._gemm(t.ptr, t.ld, t.n, t.c, Wd, W.shape[1], None, out.data_ptr(), W.shape[1])
            return _T(out, t.n, 1, 1, W.shape[1])
        if op == "Softmax":
            t = self._mat(env[node.inputs[0]], dense=True)
            assert t.h * t.w == 1 and node.attrs.get("axis", -1) in (-1, 1) and t.ld == t.c
            out = self._new(t.n, t.c)
            self.launches += 1
            _lib.check_op(self.lib.rdb_op_softmax_rows(self.device, t.ptr, t.n, t.c, out.data_ptr(), self._st()))
            return _T(out, t.n, 1, 1, t.c)
        raise NotImplementedError(f"ONNX op {op} is not supported by the B200 executor ({node})")
