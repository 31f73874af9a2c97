"""Table path: the numeric pre/post-processing around the table networks.

Reference (under rapid_doc/model/):
  T2  PaddleCls / QanythingCls preprocessing + softmax vote      table/rapid_table_self/table_cls/main.py:46-187
  T3  TablePreprocess (488 pad)                                  table/rapid_table_self/table_structure/pp_structure/pre_process.py:10-60
  T4/T6  SLANet_plus / UNET / cls networks                       ONNX files downloaded at first run — weights unavailable offline;
                                                                 the classes below take the reference's InferSession-protocol object
  T5  TableLabelDecode                                           .../pp_structure/post_process.py:11-131
  T8  RapidOrientationModel.predict's portrait / vertical-box rule   orientation/rapid_orientation_model.py:12-53

T5's reduction over the structure vocabulary ([B,T,50] probabilities -> argmax id + its probability per step) runs on the GPU
(rdb_argmax_rows, first maximum wins as np.argmax) — with a device-resident `struct_probs` only 8 bytes per step come back
instead of 200; the token walk (eos stop, <td> boxes, score mean) is short host code on those ids.
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
