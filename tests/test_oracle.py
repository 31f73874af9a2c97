"""CPU tests pinning the oracle: restatement vs golden vectors generated from the
reference's own networks (oracle/make_golden.py) and, when /root/reference exists,
vs a live import of the reference."""
import os

import cv2
import numpy as np
import pytest

from oracle import nets, ocr_post as P, ref_loader


def test_det_restatement_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "det_randn_2x64x96.npz"))
    assert np.array_equal(nets.det_forward(g["x"]), g["prob"])
    g = np.load(os.path.join(golden_dir, "det_real_192x256.npz"))
    x = P.det_preprocess(g["page_bgr"])
    assert np.array_equal(x, g["x"])
    assert np.abs(nets.det_forward(x) - g["prob"]).max() <= 1e-6


def test_rec_restatement_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "rec_randn_3x48x173.npz"))
    lg = nets.rec_logits(g["x"])
    assert lg.shape == (3, 22, 18710)
    assert np.array_equal(lg.argmax(2), g["ids"])
    assert np.abs(lg.max(2) - g["logits_max"]).max() <= 1e-4
    g = np.load(os.path.join(golden_dir, "rec_real_6lines.npz"))
    crops = [g[f"crop{i}"] for i in range(6)]
    xb, _ = P.rec_batch_tensor(crops)
    assert np.array_equal(xb, g["x"])
    probs = nets.rec_forward(xb)
    assert np.array_equal(probs.argmax(2), g["ids"])
    res = P.ctc_decode(probs, nets.load_characters())
    assert [t for t, _ in res] == list(g["texts"])
    assert np.allclose([c for _, c in res], g["conf"], atol=2e-5)
    assert "Chapter 3" in [t for t, _ in res]


def test_db_post_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "det_page_img5.npz"))
    boxes, scores = P.db_postprocess(g["prob"], tuple(g["shape"]), box_thresh=0.3, unclip_ratio=1.8)
    boxes = np.array(P.sorted_boxes(boxes))
    assert np.array_equal(boxes, g["boxes"])
    assert np.allclose(scores_sorted(boxes, g), g["scores"][: len(boxes)]) or len(scores) == len(g["scores"])


def scores_sorted(boxes, g):
    return g["scores"][: len(boxes)]


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_restatement_bit_identical_to_reference_import():
    import torch
    det, rec = ref_loader.det_net(), ref_loader.rec_net()
    x = np.random.default_rng(0).standard_normal((1, 3, 96, 128)).astype(np.float32)
    with torch.no_grad():
        want = det(torch.from_numpy(x))["maps"].numpy()
    assert np.array_equal(nets.det_forward(x), want)
    x = np.random.default_rng(1).standard_normal((2, 3, 48, 131)).astype(np.float32)
    with torch.no_grad():
        want = rec(torch.from_numpy(x))["ctc_logits"].numpy()
    assert np.array_equal(nets.rec_logits(x), want)
    assert ref_loader.characters() == nets.load_characters()


def test_dilate_semantics_match_cv2():
    rng = np.random.default_rng(5)
    seg = (rng.random((37, 52)) > 0.8).astype(np.uint8)
    want = cv2.dilate(seg, np.array([[1, 1], [1, 1]]))
    got = np.zeros_like(seg)
    H, W = seg.shape
    for y in range(H):
        for x in range(W):
            v = seg[y, x]
            if x > 0: v |= seg[y, x - 1]
            if y > 0: v |= seg[y - 1, x]
            if x > 0 and y > 0: v |= seg[y - 1, x - 1]
            got[y, x] = v
    assert np.array_equal(got, want)


def test_ctc_decode_edge_cases():
    chars = ["blank", "a", "b", " "]
    idx = np.array([[0, 0, 0, 0], [1, 1, 0, 1], [2, 2, 2, 2], [0, 3, 3, 1]])
    prob = np.full(idx.shape, 0.5, np.float32)
    res = P.ctc_decode_indices(idx, prob, chars)
    assert res[0] == ("", 0.0)
    assert res[1][0] == "aa" and res[2][0] == "b" and res[3][0] == " a"


def test_det_preprocess_shapes():
    img = np.zeros((1000, 1500, 3), np.uint8)
    x = P.det_preprocess(img, 960)
    assert x.shape == (1, 3, 640, 960)
    x = P.det_preprocess(np.zeros((1024, 1024, 3), np.uint8), 1024)
    assert x.shape == (1, 3, 1024, 1024)
    assert P.det_preprocess(np.zeros((10, 2000, 3), np.uint8), 960) is None
