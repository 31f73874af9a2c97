"""Table path: the table-structure model RapidDoc ships and the numeric pre/post-processing
around the table networks.

Reference (under rapid_doc/model/):
  T2  PaddleCls / QanythingCls preprocessing + softmax vote      table/rapid_table_self/table_cls/main.py:46-187
  T3  TablePreprocess (488 pad)                                  table/rapid_table_self/table_structure/pp_structure/pre_process.py:10-60
                                                                 (`device_batch`: normalise + pad + NHWC packing on the GPU, bit-identical)
  T4  SLANet (`slanet-1m.onnx`, ModelType.SLANET1M)              `SlaNetSession` = OrtInferSession.__call__ for this file,
                                                                 inference_engine/onnxruntime/main.py:70-76; `B200TableStructurer` =
                                                                 PPTableStructurer, table_structure/pp_structure/main.py:25-51
      SLANet_plus / UNET / cls networks                          ONNX files downloaded at first run — weights unavailable offline
  T5  TableLabelDecode                                           .../pp_structure/post_process.py:11-131
  T8  RapidOrientationModel.predict's portrait / vertical-box rule   orientation/rapid_orientation_model.py:12-53 (classifier: orientation.py)

T5's reduction over the structure vocabulary ([B,T,V] probabilities -> argmax id + its probability per step) runs on the GPU
(rdb_argmax_rows, first maximum wins as np.argmax) — with a device-resident `struct_probs` only 8 bytes per step come back
instead of 4 V; the token walk (eos stop, <td> boxes, score mean) is short host code on those ids.
"""
import cv2
import numpy as np

from . import _lib

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406])
IMAGENET_STD = np.array([0.229, 0.224, 0.225])


class TablePreprocess:
    """T3: long side -> max_len keeping the ratio (cv2.resize, INTER_LINEAR), ImageNet normalise, zero-pad to max_len^2, CHW."""

    def __init__(self, max_len=488):
        self.max_len = max_len
        self._stage, self._lut, self._lut_dev = {}, None, {}

    def __call__(self, img_list):
        if isinstance(img_list, np.ndarray):
            img_list = [img_list]
        imgs, shapes = [], []
        for img in img_list:
            if img is None:
                continue
            h, w = img.shape[:2]
            ratio = self.max_len / (max(h, w) * 1.0)
            rh, rw = int(h * ratio), int(w * ratio)
            x = cv2.resize(img, (rw, rh))
            x = (x.astype("float32") * (1 / 255.0) - IMAGENET_MEAN) / IMAGENET_STD
            pad = np.zeros((self.max_len, self.max_len, 3), dtype=np.float32)
            pad[:rh, :rw, :] = x
            imgs.append(pad.transpose((2, 0, 1)))
            shapes.append([h, w, ratio, ratio, self.max_len, self.max_len])
        return imgs, np.array(shapes)

    def device_batch(self, img_list, device=0):
        """The same preprocessing with only the uint8 resize on the host (cv2, threaded): the resized images are packed into one
        pinned uint8 canvas batch, uploaded once, and normalised / padded on the device by table lookup (rdb_op_lut_u8_nhwc4; the
        256-entry table per channel is produced by the numpy expression of __call__, so every float equals the host path's).
        -> (CUDA tensor [B, max_len, max_len, 4] fp32 NHWC, shapes as __call__)"""
        import torch
        from .dbpost import pool
        L = self.max_len
        imgs = [im for im in ([img_list] if isinstance(img_list, np.ndarray) else img_list) if im is not None]
        B = len(imgs)
        stage = self._stage.get(B)
        if stage is None:
            stage = self._stage[B] = (torch.zeros((B, L, L, 3), dtype=torch.uint8).pin_memory(), torch.zeros((B, 2), dtype=torch.int32).pin_memory())
        canvas, valid = stage
        cnp, vnp = canvas.numpy(), valid.numpy()
        shapes = []

        def one(i):
            h, w = imgs[i].shape[:2]
            ratio = L / (max(h, w) * 1.0)
            rh, rw = int(h * ratio), int(w * ratio)
            cnp[i, :rh, :rw] = cv2.resize(imgs[i], (rw, rh))
            vnp[i] = (rh, rw)
            return [h, w, ratio, ratio, L, L]
        shapes = list(pool().map(one, range(B)))
        if self._lut is None:
            v = np.arange(256, dtype=np.uint8).reshape(256, 1, 1).repeat(3, axis=2)              # [256,1,3]: value v in every channel
            lut = np.zeros((256, 1, 3), np.float32)
            lut[:] = (v.astype("float32") * (1 / 255.0) - IMAGENET_MEAN) / IMAGENET_STD          # the expression of __call__
            self._lut = np.ascontiguousarray(lut[:, 0, :].T)                                     # [3][256]
        dev = torch.device("cuda", int(device))
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream or None
            if device not in self._lut_dev:
                self._lut_dev[device] = torch.from_numpy(self._lut).to(dev)
            d_img, d_valid = canvas.to(dev, non_blocking=True), valid.to(dev, non_blocking=True)
            out = torch.empty((B, L, L, 4), dtype=torch.float32, device=dev)
            _lib.check_op(_lib.load().rdb_op_lut_u8_nhwc4(int(device), d_img.data_ptr(), d_valid.data_ptr(), self._lut_dev[device].data_ptr(), B, L, L,
                                                          out.data_ptr(), st))
            torch.cuda.current_stream(dev).synchronize()          # the pinned canvases are reused by the next call
        return out, np.array(shapes)


def cls_preprocess_paddle(imgs, resize_short=256, size=224):
    """T2 PaddleCls.batch_preprocess: short side -> 256 (LANCZOS4), centre crop 224, /255, ImageNet norm (float32), CHW."""
    mean, std = IMAGENET_MEAN.astype(np.float32), IMAGENET_STD.astype(np.float32)
    out = []
    for img in imgs:
        ih, iw = img.shape[:2]
        p = float(resize_short) / min(iw, ih)
        img = cv2.resize(img, dsize=(int(round(iw * p)), int(round(ih * p))), interpolation=cv2.INTER_LANCZOS4)
        ih, iw = img.shape[:2]
        ws, hs = (iw - size) // 2, (ih - size) // 2
        x = np.array(img[hs:hs + size, ws:ws + size, :], dtype=np.float32) / 255.0
        x -= mean
        x /= std
        out.append(x.transpose(2, 0, 1))
    return np.stack(out, axis=0).astype(dtype=np.float32, copy=False)


def cls_preprocess_q(imgs, size=224):
    """T2 QanythingCls.preprocess: gray (of the channel-swapped image) x3, PIL resize 224 (bicubic is PIL's default), norm, CHW."""
    from PIL import Image
    mean, std = IMAGENET_MEAN.astype(np.float32), IMAGENET_STD.astype(np.float32)
    out = []
    for img in imgs:
        g = cv2.cvtColor(cv2.cvtColor(img.copy(), cv2.COLOR_BGR2RGB), cv2.COLOR_BGR2GRAY)
        g = Image.fromarray(np.uint8(np.stack((g,) * 3, axis=-1))).resize((size, size))
        x = np.array(g, dtype=np.float32) / 255.0
        x -= mean
        x /= std
        out.append(x.transpose(2, 0, 1))
    return np.stack(out, axis=0).astype(np.float32)


def cls_scores(logits, names=("wired", "wireless")):
    """predict_with_scores: softmax, argmax label, max probability."""
    p = np.exp(logits - np.max(logits, axis=1, keepdims=True))
    p /= np.sum(p, axis=1, keepdims=True)
    return [names[int(i)] for i in np.argmax(p, axis=1).tolist()], np.max(p, axis=1).astype(float).tolist()


def cls_vote(cla1, score1, cla2, score2):
    """TableCls.__call__'s merge of the two classifiers: agree -> that label, disagree -> wireless; score = the lower one."""
    return [a if a == b else "wireless" for a, b in zip(cla1, cla2)], [min(a, b) for a, b in zip(score1, score2)]


class TableLabelDecode:
    """T5.  dict_character: the structure vocabulary of the model file (ONNX metadata in the reference)."""

    def __init__(self, dict_character, slanet_plus=True, merge_no_span_structure=True, device=0):
        chars = list(dict_character)
        if merge_no_span_structure:
            if "<td></td>" not in chars:
                chars.append("<td></td>")
            if "<td>" in chars:
                chars.remove("<td>")
        self.character = ["sos"] + chars + ["eos"]
        self.char_to_index = {c: i for i, c in enumerate(self.character)}
        self.td_token = ["<td>", "<td", "<td></td>"]
        self.slanet_plus, self.device = slanet_plus, device

    def argmax(self, structure_probs):
        """[B,T,V] f32 (numpy or device tensor) -> (idx [B,T] int32, prob [B,T] f32) host arrays, on the GPU."""
        B, T, V = structure_probs.shape
        idx = np.zeros((B, T), np.int32)
        val = np.zeros((B, T), np.float32)
        x = np.ascontiguousarray(structure_probs, np.float32) if isinstance(structure_probs, np.ndarray) else structure_probs.contiguous()
        if isinstance(x, np.ndarray):
            _lib.check(_lib.load().rdb_argmax_rows(int(self.device), _lib.ptr(x), B * T, V, _lib.ptr(idx), _lib.ptr(val), None))
        else:
            import torch
            di = torch.empty((B, T), dtype=torch.int32, device=x.device)
            dv = torch.empty((B, T), dtype=torch.float32, device=x.device)
            _lib.check(_lib.load().rdb_argmax_rows(int(self.device), _lib.ptr(x), B * T, V, _lib.ptr(di), _lib.ptr(dv),
                                                   torch.cuda.current_stream(x.device).cuda_stream or None))
            idx, val = di.cpu().numpy(), dv.cpu().numpy()
        return idx, val

    def __call__(self, bbox_preds, structure_probs, shape_list, ori_imgs):
        end_idx = self.char_to_index["eos"]
        ignored = (self.char_to_index["sos"], end_idx)
        idx, prob = self.argmax(structure_probs)
        bbox_preds = np.asarray(bbox_preds)
        structs, cells = [], []
        for b in range(len(idx)):
            tokens, boxes, scores = [], [], []
            for t in range(idx.shape[1]):
                c = int(idx[b][t])
                if t > 0 and c == end_idx:
                    break
                if c in ignored:
                    continue
                text = self.character[c]
                if text in self.td_token:
                    bb = bbox_preds[b, t]
                    h, w = shape_list[b][:2]
                    bb[0::2] *= w
                    bb[1::2] *= h
                    boxes.append(bb)
                tokens.append(text)
                scores.append(prob[b, t])
            cb = np.array(boxes)
            if self.slanet_plus and cb.size:
                h, w = ori_imgs[b].shape[:2]
                ratio = min(488 / h, 488 / w)
                cb[:, 0::2] *= 488 / (w * ratio)
                cb[:, 1::2] *= 488 / (h * ratio)
            if cb.size:
                cb = cb[~np.all(cb == 0, axis=1)]
            cells.append(cb)
            structs.append((["<html>", "<body>", "<table>"] + tokens + ["</table>", "</body>", "</html>"], float(np.mean(scores))))
        return structs, cells


def needs_orientation_cls(img_shape, det_res):
    """T8: the rule that decides whether RapidOrientationModel.predict runs its 4-way classifier at all
    (rapid_orientation_model.py:12-53): portrait crop (h/w > 1.2) and, when text boxes are given, at least 28 % (and >= 3) of
    them taller than wide (w/h < 0.8)."""
    h, w = img_shape[:2]
    if not ((h / w if w > 0 else 1.0) > 1.2):
        return False
    if not det_res:
        return True
    vertical = 0
    for p1, _p2, p3, _p4 in det_res:
        bw, bh = p3[0] - p1[0], p3[1] - p1[1]
        if (bw / bh if bh > 0 else 1.0) < 0.8:
            vertical += 1
    return vertical >= len(det_res) * 0.28 and vertical >= 3


class SlaNetSession:
    """T4: the SLANet ONNX graph on the B200 — what `OrtInferSession.__call__` returns for it
    (rapid_table_self/inference_engine/onnxruntime/main.py:70-76): (bbox_preds [B,T,loc], struct_probs [B,T,classes]).
    Backbone + CSP-PAN (77 convolutions) run through `onnx_run.OnnxCnn`; the SLAHead `Loop` (GRU attention, <= 501 sequential
    steps) is ONE persistent kernel launch (rdb_sla_decode, csrc/table.cu).  Tensor names are those of the Paddle2ONNX export
    (slanet-1m.onnx; SLANet_plus is exported by the same tool from the same head class)."""
    MAX_BATCH = 128          # one CTA per image, all co-resident (148 SMs)

    def __init__(self, model_path, device=0, precision="fp32"):
        import ctypes
        import torch
        from .onnx_run import OnnxCnn
        self.torch, self.device = torch, int(device)
        self.net = OnnxCnn(model_path, device, precision=precision)
        self.lib = self.net.lib
        g, c = self.net.graph, self.net.consts
        loop = [n for n in g.nodes if n.op == "Loop"]
        assert len(loop) == 1, "not a SLANet export: no Loop node"
        body = loop[0].attrs["body"]
        bconst = {n.outputs[0]: np.asarray(n.attrs["value"]) for n in body.nodes if n.op == "Constant"}
        prod = {o: n for n in g.nodes for o in n.outputs}
        feat_t = [i for i in loop[0].inputs if i in prod and prod[i].op == "Transpose"]
        assert len(feat_t) == 1
        self.feature_name = prod[prod[feat_t[0]].inputs[0]].inputs[0]            # Transpose <- Reshape <- neck output
        onehot = [n for n in body.nodes if n.op == "OneHot"][0]
        self.classes = int(bconst[onehot.inputs[1]].reshape(-1)[0])
        eq = [n for n in body.nodes if n.op == "Equal"][0]
        self.eos = int([bconst[i] for i in eq.inputs if i in bconst][0].reshape(-1)[0])
        self.max_steps = int(np.asarray(c["assign_0.tmp_0.0"]).reshape(-1)[0]) + 1
        dev = self.net.dev

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev)
        self.Wi_t = up(c["linear_0.w_0"].T)                                       # [hidden, C] = the [N, K] layout rdb_op_gemm takes
        self.hidden, self.C = self.Wi_t.shape
        self.loc_dim = int(c["linear_6.w_0"].shape[1])
        assert c["gru_cell_0.w_0"].shape == (3 * self.hidden, self.C + self.classes) and c["linear_4.w_0"].shape[1] == self.classes
        self._w = dict(Wh=up(c["linear_1.w_0"]), bh=up(c["linear_1.b_0"]), ws=up(c["linear_2.w_0"].reshape(-1)),
                       WihT=up(c["gru_cell_0.w_0"].T), WhhT=up(c["gru_cell_0.w_1"].T), bih=up(c["gru_cell_0.b_0"]), bhh=up(c["gru_cell_0.b_1"]),
                       W3=up(c["linear_3.w_0"]), b3=up(c["linear_3.b_0"]), W4=up(c["linear_4.w_0"]), b4=up(c["linear_4.b_0"]),
                       W5=up(c["linear_5.w_0"]), b5=up(c["linear_5.b_0"]), W6=up(c["linear_6.w_0"]), b6=up(c["linear_6.b_0"]))

        class W(ctypes.Structure):
            _fields_ = [("hidden", ctypes.c_int)] + [(k, ctypes.c_void_p) for k in
                                                     ("Wh", "bh", "ws", "WihT", "WhhT", "bih", "bhh", "W3", "b3", "W4", "b4", "W5", "b5", "W6", "b6")]
        self._wstruct = W(self.hidden, *[self._w[k].data_ptr() for k in ("Wh", "bh", "ws", "WihT", "WhhT", "bih", "bhh", "W3", "b3", "W4", "b4", "W5", "b5", "W6", "b6")])
        self._ctypes = ctypes
        self.launches = 0
        self.last_steps = 0

    def get_character_list(self, key="character"):
        return self.net.meta[key].splitlines()

    def __call__(self, imgs, nhwc4=None, device_out=False):
        """imgs [B,3,H,W] float32 -> (bbox_preds [B,T,loc], struct_probs [B,T,classes]) numpy, T as the graph's final Slice.
        nhwc4=(H, W): imgs is a CUDA tensor [B,H,W,4] already normalised (TablePreprocess.device_batch).  device_out: the two
        results stay on the device (torch tensors)."""
        torch, ct = self.torch, self._ctypes
        if not torch.is_tensor(imgs):                       # a CUDA tensor (already preprocessed, resident) is taken as it is
            imgs = np.asarray(imgs, np.float32)
        outs = []
        for b0 in range(0, len(imgs), self.MAX_BATCH):
            x = imgs[b0:b0 + self.MAX_BATCH]
            with torch.cuda.device(self.net.dev):
                l0 = self.net.launches
                if nhwc4 is not None:
                    feat, n, h, w, c = self.net.features(x.contiguous(), self.feature_name, nhwc4=(len(x), nhwc4[0], nhwc4[1]))
                else:
                    feat, n, h, w, c = self.net.features(x, self.feature_name)
                assert c == self.C
                st = torch.cuda.current_stream(self.net.dev).cuda_stream or None
                dev, hw, S, V, L = self.net.dev, h * w, self.max_steps, self.classes, self.loc_dim
                proj = torch.empty((n * hw, self.hidden), dtype=torch.float32, device=dev)
                _lib.check_op(self.lib.rdb_op_gemm(self.device, self.net.gemm_prec, feat.data_ptr(), c, n * hw, c, self.Wi_t.data_ptr(), self.hidden, None, 0, None, 0,
                                                   proj.data_ptr(), self.hidden, 0, st, None, 0))
                logits = torch.zeros((n, S, V), dtype=torch.float32, device=dev)
                probs = torch.empty((n, S, V), dtype=torch.float32, device=dev)
                loc = torch.zeros((n, S, L), dtype=torch.float32, device=dev)
                ids = torch.zeros((n, S), dtype=torch.int32, device=dev)
                sync = torch.zeros(2 + n + 1, dtype=torch.int32, device=dev)
                rc = self.lib.rdb_sla_decode(self.device, feat.data_ptr(), proj.data_ptr(), n, hw, c, ct.byref(self._wstruct), V, L, S, self.eos, logits.data_ptr(),
                                             probs.data_ptr(), loc.data_ptr(), ids.data_ptr(), sync.data_ptr(), sync[2:].data_ptr(), sync[2 + n:].data_ptr(), st)
                if rc < 0:
                    raise _lib.B200Error(f"rdb_sla_decode: {self.lib.rdb_sla_last_error().decode('utf-8', 'replace')}")
                self.launches += self.net.launches - l0 + 3
                total = int(sync[2 + n].item())
                self.last_steps = total
                T = min(total + 1, S)
                outs.append((loc[:, :T], probs[:, :T]) if device_out else (loc[:, :T].cpu().numpy(), probs[:, :T].cpu().numpy()))
        if len(outs) == 1:
            return outs[0]
        if device_out:
            outs = [(o[0].cpu().numpy(), o[1].cpu().numpy()) for o in outs]
        T = max(o[0].shape[1] for o in outs)                 # batches decoded separately stop at their own step: pad with untouched rows

        def pad(a, fill):
            return np.concatenate([a, np.full((a.shape[0], T - a.shape[1], a.shape[2]), fill, np.float32)], 1) if a.shape[1] < T else a
        return (np.concatenate([pad(o[0], 0.0) for o in outs]), np.concatenate([pad(o[1], 1.0 / self.classes) for o in outs]))


class B200TableStructurer:
    """Mirror of `PPTableStructurer` (table_structure/pp_structure/main.py:25-51) for ModelType.SLANET1M (the weights RapidDoc
    ships, rapid_table_self/main.py:38-40) — preprocess (T3) -> SLANet (T4) -> TableLabelDecode (T5)."""

    def __init__(self, model_path=None, model_type="slanet_1m", device=0, precision="fp32"):
        import os
        from .weights import WEIGHTS_DIR
        if model_path is None:
            model_path = os.path.join(WEIGHTS_DIR, "slanet-1m.onnx")
        self.session = SlaNetSession(model_path, device, precision=precision)
        self.preprocess_op = TablePreprocess()
        self.postprocess_op = TableLabelDecode(self.session.get_character_list(), slanet_plus=(model_type == "slanet_plus"), device=device)

    def __call__(self, ori_imgs):
        """Host crops in, (structure tokens + score, cell boxes) out.  Only the uint8 resize runs on the host; normalisation,
        network, decode loop and the arg-max of the label decode run on the device."""
        x, shape_lists = self.preprocess_op.device_batch(ori_imgs, self.session.device)
        L = self.preprocess_op.max_len
        bbox_preds, struct_probs = self.session(x, nhwc4=(L, L), device_out=len(x) <= self.session.MAX_BATCH)
        if not isinstance(bbox_preds, np.ndarray):
            bbox_preds = bbox_preds.cpu().numpy()
        return self.postprocess_op(bbox_preds, struct_probs, shape_lists, ori_imgs)
