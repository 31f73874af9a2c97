"""CPU tests of the product's host-side logic (rapiddoc_b200/ocr.py) against the oracle
restatement: DB post-process on a golden prob map, rec crop packing, box sorting."""
import os

import numpy as np

from oracle import ocr_post as P
from rapiddoc_b200 import ocr as O


def test_db_postprocess_host_matches_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "det_page_img5.npz"))
    prob = g["prob"][0, 0]
    shape = tuple(int(v) for v in g["shape"])
    for box_thresh, ratio, dil in [(0.3, 1.8, True), (0.5, 1.6, True), (0.3, 1.8, False)]:
        bitmap = P.db_bitmap(prob, 0.3, dil)
        post = O.DBPostProcess(0.3, box_thresh, 1000, ratio, dil)
        boxes, scores = post(prob, bitmap, shape)
        want_b, want_s = P.db_postprocess(g["prob"], shape, 0.3, box_thresh, ratio, dil)
        assert np.array_equal(np.asarray(boxes), np.asarray(want_b))
        assert np.allclose(scores, want_s, rtol=0, atol=0)
        assert [b.tolist() for b in O.sorted_boxes(boxes)] == [b.tolist() for b in P.sorted_boxes(want_b)]
    assert len(boxes) >= 10


def test_rec_pack_equals_resize_norm_img(golden_dir):
    g = np.load(os.path.join(golden_dir, "rec_real_6lines.npz"))
    crops = [g[f"crop{i}"] for i in range(6)]
    rec = O.B200TextRecognizer.__new__(O.B200TextRecognizer)
    rec.rec_image_shape = [3, 48, 320]
    ratio = max([320 / 48] + [c.shape[1] / c.shape[0] for c in crops])
    buf, vw = rec._pack(crops, ratio)
    x = ((buf.astype(np.float32) / 255 - 0.5) / 0.5).transpose(0, 3, 1, 2)
    for i in range(6):
        x[i, :, :, vw[i]:] = 0
    want, r2 = P.rec_batch_tensor(crops)
    assert r2 == ratio and x.shape == want.shape
    assert np.array_equal(x, want)


def test_rotate_crop_shapes():
    img = (np.random.default_rng(0).random((200, 300, 3)) * 255).astype(np.uint8)
    quad = np.float32([[10, 20], [110, 22], [109, 52], [9, 50]])
    c = O.get_rotate_crop_image(img, quad)
    assert c.shape[0] in (29, 30, 31) and c.shape[1] in (99, 100, 101)
    tall = np.float32([[10, 10], [30, 10], [30, 150], [10, 150]])
    c = O.get_rotate_crop_image(img, tall)
    assert c.shape[1] > c.shape[0]          # rotated by 90 degrees (h/w >= 2)


def test_get_word_info_grouping():
    """ocr_patch.py:333-389 semantics: CN/EN runs, spaces are their own word, gaps > 5 columns split a word."""
    sel = np.zeros(40, bool)
    sel[[2, 4, 6, 9, 15, 17, 19]] = True
    wi = O.get_word_info("ab c邹12", sel)
    assert wi.words == [["a", "b"], [" "], ["c"], ["邹"], ["1", "2"]]
    assert wi.word_cols == [[2, 4], [6], [9], [15], [17, 19]]
    assert len(wi.word_types) == 5 and wi.word_types[3] != wi.word_types[0]
    sel = np.zeros(30, bool)
    sel[[1, 2, 12, 13]] = True                      # 10-column gap inside an EN run -> two words
    wi = O.get_word_info("abcd", sel)
    assert wi.words == [["a", "b"], ["c", "d"]]
    assert O.get_word_info("", np.zeros(5, bool)).words == []


def test_bench_roofline_model_knows_every_fused_kernel():
    """bench.py's algorithmic-bytes model (roofline.achieved) must cover the launch names the engines emit."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = bench.WORKLOADS["det"]
    P = 32 * 1024 * 1024 // 16
    assert bench.algorithmic_bytes(f"stem_planar[P={P}]", 2, wl, 32) == P * (4 * 24 + 48) * 2
    assert bench.algorithmic_bytes(f"head_planar[P={P}]", 2, wl, 32) > P * 16 * 5
    assert bench.algorithmic_bytes(f"mlp_tc[M={P},C=48,N=48,res=1]", 2, wl, 32) == P * (48 + 48 + 48) * 2
    assert bench.algorithmic_bytes(f"stem1_tc[P={4 * P},N=24]", 2, wl, 32) == 4 * P * (12 + 48)
    assert bench.algorithmic_bytes(f"dwconv7x7_h2[P={P},C=96,s=1]", 2, wl, 32) == P * 96 * 2 * 2
    assert bench.algorithmic_bytes("se_fc", 2, wl, 32) is None       # latency-bound helper: no byte model, never the roofline kernel


def test_crop_geometry_batch_equals_per_box_and_translation_shortcut_is_exact():
    """crop_geometry_batch: sizes / rot90 decisions identical to the per-box reference arithmetic; for axis-aligned integer
    boxes the exact translation it substitutes for cv2's getPerspectiveTransform + invert yields the very same warp
    (checked through the cv2-pinned CPU restatement of warpPerspective and against cv2 itself)."""
    import cv2
    from oracle import warp
    from rapiddoc_b200.ocr import crop_geometry, crop_geometry_batch
    rng = np.random.default_rng(3)
    page = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    boxes = []
    for t in range(80):
        x0, y0 = int(rng.integers(0, 300)), int(rng.integers(0, 250))
        w, h = int(rng.integers(1, 100)), int(rng.integers(1, 50))
        if t % 4 == 0:       # rotated / skewed integer quad
            boxes.append([[x0, y0 + 3], [x0 + w, y0], [x0 + w + 2, y0 + h], [x0 + 1, y0 + h + 4]])
        elif t % 9 == 0:     # tall box -> rot90
            boxes.append([[x0, y0], [x0 + 8, y0], [x0 + 8, y0 + 40], [x0, y0 + 40]])
        else:
            boxes.append([[x0, y0], [x0 + w, y0], [x0 + w, y0 + h], [x0, y0 + h]])
    boxes.append([[5, 5], [5, 5], [5, 5], [5, 5]])        # degenerate
    keep, sizes, minv, rot = crop_geometry_batch(boxes)
    per = [crop_geometry(b) for b in boxes]
    assert list(keep) == [i for i, g in enumerate(per) if g is not None]
    for k, i in enumerate(keep):
        cw, ch, mi, r = per[i]
        assert (int(sizes[k][0]), int(sizes[k][1]), int(rot[k])) == (cw, ch, r)
        assert np.abs(minv[k] - mi.reshape(9)).max() <= 1e-9
        a = warp.warp_perspective_cubic_replicate(page, minv[k].reshape(3, 3), (cw, ch))
        pts = np.float32(boxes[i])
        M = cv2.getPerspectiveTransform(pts, np.float32([[0, 0], [cw, 0], [cw, ch], [0, ch]]))
        want = cv2.warpPerspective(page, M, (cw, ch), borderMode=cv2.BORDER_REPLICATE, flags=cv2.INTER_CUBIC)
        assert np.array_equal(a, want)
    # non-integer coordinates fall back to the per-box cv2 path
    fb = [[[1.5, 2.25], [60.5, 2.0], [60.0, 20.5], [1.0, 21.0]]]
    keep, sizes, minv, rot = crop_geometry_batch(fb)
    g = crop_geometry(fb[0])
    assert (int(sizes[0][0]), int(sizes[0][1]), int(rot[0])) == (g[0], g[1], g[3]) and np.array_equal(minv[0], g[2].reshape(9))
