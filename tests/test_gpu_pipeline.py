"""GPU tests of the window-level OCR path (round 2): device box scorer, cross-page crop + recognition batching, and the whole
det -> boxes -> crops -> rec flow of `B200OcrModel.ocr_pages` against the CPU oracle's restatement of the reference flow
(oracle/pipeline.py).  Everything goes through the C-ABI."""
import os

import cv2
import numpy as np
import pytest

from oracle import ocr_post as P
from oracle import pipeline as OP
from rapiddoc_b200 import PREC_FP16, PREC_FP32, _lib, dbpost, synth
from rapiddoc_b200.ocr import B200OcrModel, DeviceCrops, crop_geometry, get_rotate_crop_image

pytestmark = pytest.mark.gpu


def _prob_maps(n=3, h=256, w=384):
    from test_dbpost import _synthetic_prob
    return np.stack([_synthetic_prob(10 + s, h, w, 14) for s in range(n)])


def test_box_scores_equal_cv2_mean_over_fillpoly():
    import torch
    probs = _prob_maps()
    n, h, w = probs.shape
    rng = np.random.default_rng(0)
    quads, idx = [], []
    for t in range(600):
        cx, cy = rng.uniform(-5, w + 5), rng.uniform(-5, h + 5)          # some boxes stick out of the page -> cv2 fallback
        bw, bh, a = rng.uniform(4, 300), rng.uniform(3, 40), rng.uniform(-1.5, 1.5) if t % 3 == 0 else rng.uniform(-0.2, 0.2)
        quads.append(cv2.boxPoints(((cx, cy), (bw, bh), float(np.degrees(a)))))
        idx.append(int(rng.integers(0, n)))
    quads = dbpost.mini_boxes(np.stack(quads).astype(np.float32))
    idx = np.array(idx, np.int32)
    want = dbpost.cv2_score_fn(probs)(quads, idx)
    d = torch.from_numpy(probs).cuda()
    got = dbpost.gpu_score_fn(0, d, n, h, w)(np.ascontiguousarray(quads), idx)
    assert np.abs(got - want).max() <= 1e-12
    # host prob maps through the same entry
    scores = np.zeros(len(quads)); flags = np.zeros(len(quads), np.int32)
    _lib.check(_lib.load().rdb_db_box_scores(0, probs.ctypes.data, n, h, w, len(quads), quads.ctypes.data, idx.ctypes.data, scores.ctypes.data,
                                            flags.ctypes.data, None))
    ok = flags == 0
    assert ok.sum() > 200 and np.abs(scores[ok] - want[ok]).max() <= 1e-12


def test_resize_pack_slots_bit_exact_vs_cv2():
    import torch
    rng = np.random.default_rng(1)
    crops = [rng.integers(0, 256, (int(rng.integers(8, 70)), int(rng.integers(8, 500)), 3), dtype=np.uint8) for _ in range(40)]
    nbytes = np.array([c.size for c in crops], np.int64)
    offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
    src = torch.from_numpy(np.concatenate([c.ravel() for c in crops])).cuda()
    sizes = np.array([[c.shape[1], c.shape[0]] for c in crops], np.int32)
    pitch = np.array([int(rng.integers(320, 700)) for _ in crops], np.int32)
    dw = np.array([min(int(np.ceil(48 * c.shape[1] / c.shape[0])), int(p)) for c, p in zip(crops, pitch)], np.int32)
    slot = 48 * pitch.astype(np.int64) * 3
    doffs = np.concatenate([[0], np.cumsum(slot)[:-1]]).astype(np.int64)
    dst = torch.full((int(slot.sum()),), 7, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().rdb_resize_pack_slots(0, src.data_ptr(), int(src.numel()), len(crops), offs.ctypes.data, sizes.ctypes.data, dw.ctypes.data,
                                                 doffs.ctypes.data, pitch.ctypes.data, dst.data_ptr(), int(dst.numel()), 48, None))
    out = dst.cpu().numpy()
    for i, c in enumerate(crops):
        got = out[doffs[i]: doffs[i] + slot[i]].reshape(48, pitch[i], 3)
        assert np.array_equal(got[:, :dw[i]], cv2.resize(c, (int(dw[i]), 48)))
        assert not got[:, dw[i]:].any()


def test_warp_crops_batch_bit_exact_vs_cv2():
    import torch
    rng = np.random.default_rng(2)
    pages = rng.integers(0, 256, (3, 200, 300, 3), dtype=np.uint8)
    quads, pidx = [], []
    for t in range(40):
        cx, cy = rng.uniform(20, 280), rng.uniform(20, 180)
        bw, bh, a = rng.uniform(10, 200), rng.uniform(6, 40), rng.uniform(-0.4, 0.4)
        if t % 7 == 0:
            bw, bh = bh, bw * 0.8        # tall box -> rot90
        quads.append(P.order_points_clockwise(cv2.boxPoints(((cx, cy), (bw, bh), float(np.degrees(a))))))
        pidx.append(int(rng.integers(0, 3)))
    geo = [crop_geometry(q) for q in quads]
    keep = [i for i, g in enumerate(geo) if g is not None]
    minv = np.stack([geo[i][2].reshape(9) for i in keep])
    sizes = np.array([[geo[i][0], geo[i][1]] for i in keep], np.int32)
    rot = np.array([geo[i][3] for i in keep], np.int32)
    nbytes = sizes[:, 0].astype(np.int64) * sizes[:, 1] * 3
    offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
    pi = np.array([pidx[i] for i in keep], np.int32)
    d_pages = torch.from_numpy(pages).cuda()
    out = torch.empty(int(nbytes.sum()), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().rdb_warp_crops_batch(0, d_pages.data_ptr(), 3, 200, 300, len(keep), pi.ctypes.data, minv.ctypes.data, sizes.ctypes.data,
                                                rot.ctypes.data, out.data_ptr(), offs.ctypes.data, int(out.numel()), None))
    o = out.cpu().numpy()
    for k, i in enumerate(keep):
        want = get_rotate_crop_image(pages[pidx[i]], quads[i])
        got = o[offs[k]: offs[k] + nbytes[k]].reshape(want.shape)
        assert np.array_equal(got, want)


def _golden_page(golden_dir):
    g = np.load(os.path.join(golden_dir, "page_img5_e2e.npz"))
    return cv2.imdecode(g["png"], cv2.IMREAD_COLOR), g


def test_det_window_fp32_boxes_exact_and_batch_invariant(golden_dir):
    img, g = _golden_page(golden_dir)
    model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, enable_merge_det_boxes=False, precision=PREC_FP32)
    res = model.det_batch_predict([img, img.copy(), img.copy()], max_batch_size=1)
    for boxes, _ in res:
        assert np.array_equal(np.asarray(boxes, np.float32), g["boxes"])


@pytest.mark.parametrize("prec", [PREC_FP32, PREC_FP16])
def test_ocr_pages_vs_oracle_pipeline(prec):
    """Window flow on synthetic pages: boxes and texts of every line against the CPU oracle of the reference flow."""
    pages = list(synth.det_pages(3, 384, 640, seed=5, lines=8))
    want = OP.ocr_pages(pages, limit_side_len=960, box_thresh=0.3, unclip_ratio=1.8, merge=True, rec_batch_num=6)
    model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, enable_merge_det_boxes=True, precision=prec)
    got = model.ocr_pages(pages)
    lines = mism = boxbad = 0
    for gp, wp in zip(got, want):
        gp, wp = gp or [], wp or []
        assert len(gp) == len(wp)
        for (gb, (gt, gs)), (wb, (wt, ws)) in zip(gp, wp):
            lines += 1
            d = np.abs(np.asarray(gb, np.float32) - np.asarray(wb, np.float32)).max()
            boxbad += d > 0
            mism += gt != wt
            if prec == PREC_FP32:
                assert abs(gs - ws) <= 2e-4
    assert lines >= 12
    print(f"prec={prec}: lines={lines} box_mismatch={boxbad} text_mismatch={mism}")
    if prec == PREC_FP32:
        assert boxbad == 0 and mism == 0
    else:
        assert mism <= max(1, lines // 20)
    # the same pages already resident on the GPU (bench `value` path) and the host-crop flow of the reference callers
    import torch
    again = model.ocr_pages(torch.from_numpy(np.stack(pages)).cuda())
    assert [[(b, t) for b, (t, _) in (p or [])] for p in again] == [[(b, t) for b, (t, _) in (p or [])] for p in got]
    dets = model.det_batch_predict(pages, max_batch_size=1)
    # host crops -> ocr(det=False) must give the texts of the fused path
    k = 0
    boxes0 = model._post_boxes(dets[k][0], None)
    crops0 = [get_rotate_crop_image(pages[k], np.asarray(b, np.float32)) for b in boxes0]
    rec = model.ocr(crops0, det=False)[0]
    fused = {tuple(np.asarray(b, np.float32).ravel()): t for b, (t, _) in (got[k] or [])}
    if prec == PREC_FP32:
        same = sum(fused.get(tuple(np.asarray(b, np.float32).ravel())) == t for b, (t, sc) in zip(boxes0, rec) if sc >= 0.5)
        assert same >= len([1 for _, (t, sc) in zip(boxes0, rec) if sc >= 0.5]) - 1


def test_ocr_pages_stream_equals_ocr_pages():
    pages_a = list(synth.det_pages(3, 384, 640, seed=5, lines=8))
    pages_b = list(synth.det_pages(2, 256, 512, seed=6, lines=5))
    model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, precision=PREC_FP16)
    want = [model.ocr_pages(p) for p in (pages_a, pages_b, pages_a)]
    got = list(model.ocr_pages_stream([pages_a, pages_b, pages_a]))
    assert got == want
    assert list(model.ocr_pages_stream([])) == []


def test_recognizer_window_path_on_golden_lines(golden_dir):
    g = np.load(os.path.join(golden_dir, "rec_real_6lines.npz"))
    crops = [g[f"crop{i}"] for i in range(6)]
    model = B200OcrModel(precision=PREC_FP32)
    r = model.text_recognizer(crops)
    assert list(r.txts) == list(g["texts"])
    assert np.abs(np.array(r.scores) - g["conf"]).max() <= 2e-4
    # larger batches than the reference default: grouping changes the padded width, texts on these clean lines do not
    model.text_recognizer.rec_batch_num = 4
    r2 = model.text_recognizer(crops + crops)
    assert len(r2.txts) == 12


def test_fp16_real_page_decisions(golden_dir):
    """fp16 mode on the real page: report (and bound) the decision differences against the reference-net goldens."""
    img, g = _golden_page(golden_dir)
    model = B200OcrModel(det_db_box_thresh=0.3, det_db_unclip_ratio=1.8, enable_merge_det_boxes=False, precision=PREC_FP16)
    (boxes, _), = model.det_batch_predict([img], max_batch_size=1)
    assert len(boxes) == len(g["boxes"])
    d = np.abs(np.asarray(boxes, np.float32) - g["boxes"])
    res = model.ocr(img, det=True, rec=True)[0]
    texts = [r[1][0] for r in res]
    bad = sum(a != b for a, b in zip(texts, list(g["texts"])))
    print(f"fp16 real page: boxes differing {int((d.max(axis=(1, 2)) > 0).sum())}/{len(boxes)} (max {d.max():.0f} px), texts differing {bad}/{len(texts)}")
    assert d.max() <= 2 and bad <= 1


def test_pool_cache_is_bounded():
    from rapiddoc_b200.engine import DetEngine
    eng = DetEngine(device=0, precision=PREC_FP16)
    lib = _lib.load()
    _lib.check(lib.rdb_det_set_pool_cap_bytes(eng._h, 64 << 20))
    rng = np.random.default_rng(0)
    peak = 0
    for hgt, wid in [(256, 256), (320, 512), (512, 384), (640, 640), (256, 256), (96, 960)]:
        eng.infer_u8(rng.integers(0, 256, (2, hgt, wid, 3), dtype=np.uint8))
        peak = max(peak, int(lib.rdb_det_pool_bytes(eng._h)))
    # idle cache <= cap per lane (4 lanes) after every call, whatever shapes were seen
    assert int(lib.rdb_det_pool_bytes(eng._h)) <= 4 * (64 << 20) + (8 << 20), lib.rdb_det_pool_bytes(eng._h)


def test_two_devices_in_one_process():
    lib = _lib.load()
    if lib.rdb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch
    from rapiddoc_b200.multi import B200OcrPool
    pages = list(synth.det_pages(4, 256, 512, seed=3, lines=6))
    one = B200OcrModel(det_db_box_thresh=0.3, precision=PREC_FP16, device=0)
    want = one.det_batch_predict(pages)
    pool = B200OcrPool([0, 1], det_db_box_thresh=0.3, precision=PREC_FP16)
    got = pool.det_batch_predict(pages)
    for (a, _), (b, _) in zip(got, want):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    assert torch.cuda.current_device() == 0          # the C-ABI restores the caller's device
