"""The ONNX CNNs of the orientation classifier and the seal detector: reader, CPU oracle pinned to
OpenCV's DNN importer, and the CUDA executor against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_golden_onnx as MG          # noqa: E402
from oracle import onnx_ref                        # noqa: E402
from rapiddoc_b200 import onnx_lite                # noqa: E402

ORI = os.path.join(ROOT, "weights", "rapid_orientation.onnx")
SEAL = os.path.join(ROOT, "weights", "pp-ocrv4_mobile_seal_det.onnx")
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "onnx_cases.npz"))


def test_reader_sees_the_whole_graph():
    g = onnx_lite.load(ORI)
    assert g.inputs == ["x"] and g.outputs == ["fetch_name_0"] and len(g.nodes) == 115
    assert g.meta["character"].splitlines() == ["0", "90", "180", "270"]
    assert sum(n.op == "Conv" for n in g.nodes) == 32 and g.init["linear_0.w_0_deepcopy_144"].shape == (1280, 4)
    conv = g.nodes[0]
    assert conv.attrs["strides"] == [2, 2] and conv.attrs["pads"] == [1, 1, 1, 1] and conv.attrs["group"] == 1
    s = onnx_lite.load(SEAL)
    assert len(s.nodes) == 526 and sum(n.op == "ConvTranspose" for n in s.nodes) == 2
    hs = {round(n.attrs["alpha"], 4) for n in s.nodes if n.op == "HardSigmoid"}
    assert hs == {0.1667, 0.2}
    assert abs(sum(a.size for a in s.init.values()) - 1171745) < 10


def test_oracle_matches_the_cv2_dnn_golden():
    _, xs = MG.orientation_inputs()
    y = onnx_ref.run(ORI, xs)                                   # batch 4 in one run; the golden was made one image at a time
    assert np.abs(y - GOLD["orientation_scores"]).max() < 2e-6
    assert list(np.argmax(y, 1)) == [0, 3, 2, 1]                # page rotated by 0 / 90ccw / 180 / 90cw
    p = onnx_ref.run(SEAL, MG.seal_input())[0, 0]
    assert np.abs(p - GOLD["seal_prob"].astype(np.float32)).max() < 1e-3      # fp16 storage of the golden
    assert (p > 0.2).sum() > 5000


def test_oracle_matches_cv2_dnn_live():
    cv2 = pytest.importorskip("cv2")
    x = np.random.RandomState(3).randn(1, 3, 224, 224).astype(np.float32)
    net = cv2.dnn.readNetFromONNX(ORI)
    net.setInput(x)
    assert np.abs(onnx_ref.run(ORI, x) - net.forward()).max() < 2e-6
    x = MG.seal_input(seed=1, size=256)
    net = cv2.dnn.readNetFromONNX(SEAL)
    net.setInput(x)
    assert np.abs(onnx_ref.run(SEAL, x) - net.forward()).max() < 3e-4        # cv2.dnn fuses / reorders the fp32 sums


def test_orientation_label_follows_the_reference_preprocess():
    rots, _ = MG.orientation_inputs()
    labels = [onnx_ref.orientation(ORI, r)[0] for r in rots]
    assert labels == ["0", "270", "180", "90"]


def test_executor_plan_fuses_elementwise_runs():
    """Compile only (no GPU work): Identity nodes vanish, Mul(HardSigmoid(y), y) pairs become one step, Conv + BatchNormalization
    + activation (and Conv + bias Add + Relu) collapse into the conv launch."""
    from rapiddoc_b200.onnx_run import OnnxCnn
    ops = [n.op for n in OnnxCnn(SEAL, compile_only=True).nodes]
    assert "Identity" not in ops and ops.count("HardSwishAB") >= 20
    assert ops.count("HardSigmoid") == 34 - ops.count("HardSwishAB")          # the rest are squeeze-excite gates
    ori = OnnxCnn(ORI, compile_only=True).nodes
    assert len(ori) < 50 and not any(n.op in ("BatchNormalization", "HardSwish", "Relu") for n in ori)     # 115 graph nodes
    convs = [n for n in ori if n.op == "Conv"]
    assert sum("fold_bn" in n.attrs and n.attrs.get("fold_act") == 6 for n in convs) == 27
    sla = OnnxCnn(os.path.join(ROOT, "weights", "slanet-1m.onnx"), compile_only=True).nodes
    assert sum(n.op == "Conv" and "fold_bn" in n.attrs and n.attrs.get("fold_act") == 6 for n in sla) == 73


def test_seal_polygons_from_the_oracle_probability_map():
    """Polygon-mode DB post-process on the seal map (restated from the published algorithm, parity unpinned: rapidocr absent):
    the curved text band and the straight line come out as two clean (loop-free) polygons around their text."""
    cv2 = pytest.importorskip("cv2")
    from rapiddoc_b200.orientation import polygons_from_bitmap, sort_poly_boxes
    p = onnx_ref.run(SEAL, MG.seal_input())[0, 0]
    polys, scores = polygons_from_bitmap(p, (p > 0.2).astype(np.uint8), 320, 320)
    assert len(polys) == 2 and all(s > 0.6 for s in scores)
    polys = sort_poly_boxes(polys)
    arc, line = polys
    assert arc[:, 1].min() < 40 and arc[:, 0].max() > 280 and cv2.contourArea(arc.astype(np.float32)) > 10000       # the band along the ring
    assert 130 < line[:, 1].min() and line[:, 1].max() < 180 and line[:, 0].max() - line[:, 0].min() > 80           # "CONTRACT SEAL"
    for q in polys:                                          # no self-intersection loops: filled area == shoelace area
        m = np.zeros((320, 320), np.uint8)
        cv2.fillPoly(m, [q.reshape(-1, 1, 2)], 1)
        assert abs(int(m.sum()) - cv2.contourArea(q.astype(np.float32))) < 0.08 * m.sum() + 60
    big, _ = polygons_from_bitmap(p, (p > 0.2).astype(np.uint8), 640, 960)        # scaled to the source image size
    assert max(q[:, 0].max() for q in big) > 560 and max(q[:, 1].max() for q in big) > 780


# ------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_orientation_executor_matches_the_oracle():
    from rapiddoc_b200.onnx_run import OnnxCnn
    rots, xs = MG.orientation_inputs()
    net = OnnxCnn(ORI, 0)
    y = net(xs)
    ref = onnx_ref.run(ORI, xs)
    assert y.shape == (4, 4)
    assert np.abs(y - ref).max() < 1e-4, np.abs(y - ref).max()               # fp32 both sides, different summation order
    assert np.abs(y - GOLD["orientation_scores"]).max() < 1e-4
    assert net.launches < 160                                               # 115 graph nodes -> fused launches
    y1 = net(xs[:1])                                                        # batch 1 = the reference's own call shape
    assert np.abs(y1 - ref[:1]).max() < 1e-4


@pytest.mark.gpu
def test_direct_stem_conv_is_bit_identical_to_im2col_gemm():
    """rdb_op_conv3x3_c4 (first layer, direct form) against the im2col + fp32 GEMM path it replaces: same products, same order."""
    import torch
    from rapiddoc_b200.onnx_run import OnnxCnn
    _, xs = MG.orientation_inputs()
    a, b = OnnxCnn(ORI, 0), OnnxCnn(ORI, 0)
    b.direct_stem = False
    name = a.nodes[0].outputs[0]
    fa = a.features(xs, name)[0].clone()
    fb = b.features(xs, name)[0].clone()
    assert fa.shape == fb.shape == (4 * 112 * 112, 16) and torch.equal(fa, fb)
    assert a.launches < b.launches
    assert np.array_equal(a(xs), b(xs))
    _, x, _ = MG.table_inputs()
    s, t = OnnxCnn(SLANET, 0), OnnxCnn(SLANET, 0)
    t.direct_stem = False
    assert torch.equal(s.features(x, "hardswish_72.tmp_0")[0], t.features(x, "hardswish_72.tmp_0")[0])


@pytest.mark.gpu
def test_orientation_model_interface():
    from rapiddoc_b200.orientation import B200Orientation, B200OrientationModel
    import cv2
    rots, _ = MG.orientation_inputs()
    eng = B200Orientation(device=0)
    assert [eng(r)[0] for r in rots] == ["0", "270", "180", "90"]
    assert np.abs(eng.scores(rots) - GOLD["orientation_scores"]).max() < 1e-4
    m = B200OrientationModel(device=0)
    portrait = cv2.cvtColor(rots[0], cv2.COLOR_BGR2RGB)                     # 640 x 480 portrait
    assert m.predict(portrait) == onnx_ref.orientation(ORI, portrait)[0]
    assert m.predict(cv2.cvtColor(rots[1], cv2.COLOR_BGR2RGB)) == "0"       # landscape: the classifier is not run
    flat = [[[0, 0], [100, 0], [100, 20], [0, 20]]] * 5
    assert m.predict(portrait, flat) == "0"                                 # portrait but horizontal text boxes


@pytest.mark.gpu
@pytest.mark.parametrize("size", [320, 256])
def test_seal_detector_executor_matches_the_oracle(size):
    from rapiddoc_b200.onnx_run import OnnxCnn
    x = MG.seal_input(seed=0 if size == 320 else 1, size=size)
    net = OnnxCnn(SEAL, 0)
    p = net(x)
    ref = onnx_ref.run(SEAL, x)
    assert p.shape == ref.shape == (1, 1, size, size)
    d = np.abs(p - ref)
    assert d.max() < 2e-3 and d.mean() < 1e-5, (d.max(), d.mean())          # sigmoid of fp32 sums in another order
    assert ((p > 0.2) != (ref > 0.2)).sum() <= 2                            # the bitmap at the seal threshold
    if size == 320:
        assert np.abs(p[0, 0] - GOLD["seal_prob"].astype(np.float32)).max() < 3e-3
    assert net.launches < 330                                               # 526 graph nodes


@pytest.mark.gpu
def test_seal_detector_interface():
    from rapiddoc_b200 import synth
    from rapiddoc_b200.orientation import B200SealDetector, sort_poly_boxes
    det = B200SealDetector(device=0)
    img = synth.seal_image(0, 320)
    prob, bitmap = det(img)
    assert prob.shape == (736, 736) and bitmap.dtype == np.uint8             # short side scaled up to 736 (limit_type 'min')
    ref = onnx_ref.run(SEAL, det.preprocess(img))[0, 0]
    assert ((prob > 0.2) != (ref > 0.2)).mean() < 1e-4
    polys = [np.array([[0, 30], [5, 40], [9, 31]]), np.array([[0, 3], [5, 4], [9, 9]]), np.array([[0, 13], [5, 14]])]
    assert [int(p[0, 1]) for p in sort_poly_boxes(polys)] == [3, 13, 30]
    found, scores = det.detect(img)                     # at the 736-px network scale only the curved band is seal text
    assert len(found) >= 1 and all(s > 0.6 for s in scores) and [q[:, 1].min() for q in found] == sorted(q[:, 1].min() for q in found)
    arc = found[0]
    assert arc[:, 1].min() < 60 and arc[:, 0].max() > 270 and arc[:, 0].max() <= 320 and arc[:, 1].max() <= 320     # source-image coordinates


# ------------------------------------------------------------------------------------------------------- SLANet (T4)
SLANET = os.path.join(ROOT, "weights", "slanet-1m.onnx")


def test_slanet_oracle_runs_the_loop_of_the_file():
    """The node-by-node interpreter executes the Loop / If subgraphs literally; structure of three synthetic tables."""
    imgs, x, _ = MG.table_inputs()
    loc, probs = onnx_ref.run(SLANET, x[:1])
    ids = probs.argmax(-1)[0]
    g = GOLD["slanet_ids"][0]
    T = len(ids)
    assert T == 26 and list(ids) == list(g[:T])                 # batch 1 stops at its own eos (+ the untouched row)
    assert ids[24] == 29 and ids[25] == 0 and np.allclose(probs[0, 25], 1 / 30) and np.all(loc[0, 25] == 0)
    assert np.abs(loc[0, :T - 1] - GOLD["slanet_loc"][0, :T - 1].astype(np.float32)).max() < 2e-3   # (in the batch-3 golden this row decodes on)
    chars = onnx_lite.load(SLANET).meta["character"].splitlines()
    from rapiddoc_b200 import table
    dec = table.TableLabelDecode(chars, slanet_plus=False)
    assert dec.character[29] == "eos" and dec.character[28] == "<td></td>"
    toks = [dec.character[i] for i in ids[:24]]
    assert toks.count("<tr>") == 4 and toks.count("<td></td>") == 12


def test_slanet_session_reads_the_head_from_the_graph():
    from rapiddoc_b200.table import SlaNetSession
    s = SlaNetSession.__new__(SlaNetSession)
    g = onnx_lite.load(SLANET)
    loop = [n for n in g.nodes if n.op == "Loop"][0]
    body = loop.attrs["body"]
    assert len(body.nodes) == 235 and sum(n.op == "MatMul" for n in body.nodes) == 10
    assert [n for n in body.nodes if n.op == "If"][0].attrs["then_branch"].outputs


@pytest.mark.gpu
def test_slanet_backbone_matches_the_oracle():
    from rapiddoc_b200.onnx_run import OnnxCnn
    _, x, _ = MG.table_inputs()
    net = OnnxCnn(SLANET, 0)
    feat, n, h, w, c = net.features(x, "hardswish_72.tmp_0")
    ref, = onnx_ref.run(SLANET, x, outputs=["hardswish_72.tmp_0"])
    assert (n, c, h, w) == ref.shape == (3, 96, 16, 16)
    got = feat.view(n, h, w, c).permute(0, 3, 1, 2).cpu().numpy()
    assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), np.abs(got - ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1, 3])
def test_slanet_session_matches_the_oracle(batch):
    """One persistent launch for the whole decode loop: same stop step, same tokens, probabilities / boxes to fp32 rounding.
    batch 3 = three tables with different lengths: every row keeps decoding until the LAST one has emitted eos."""
    from rapiddoc_b200.table import SlaNetSession
    _, x, _ = MG.table_inputs()
    x = x[:batch]
    s = SlaNetSession(SLANET, 0)
    loc, probs = s(x)
    rloc, rprobs = onnx_ref.run(SLANET, x)
    assert loc.shape == rloc.shape and probs.shape == rprobs.shape, (loc.shape, rloc.shape)
    assert np.array_equal(probs.argmax(-1), rprobs.argmax(-1))
    assert np.abs(probs - rprobs).max() < 2e-4 and np.abs(loc - rloc).max() < 2e-4, (np.abs(probs - rprobs).max(), np.abs(loc - rloc).max())
    if batch == 3:
        assert np.array_equal(probs.argmax(-1), GOLD["slanet_ids"]) and loc.shape[1] == 42
    assert s.launches < 250                                       # 830 graph nodes + <= 501 loop iterations


@pytest.mark.gpu
def test_slanet_batch_of_diverse_tables_matches_the_oracle():
    """Six tables from 3x2 to 8x5 cells, ruled and borderless, decoded together: one common stop step, per-row tokens and boxes
    equal to the node-by-node oracle run on the same batch."""
    sys.path.insert(0, ROOT)
    import bench
    from rapiddoc_b200.table import SlaNetSession, TablePreprocess
    imgs = [bench.table_inputs(32)[i] for i in (0, 5, 11, 17, 22, 31)]
    x, _ = TablePreprocess()(imgs)
    x = np.asarray(x, np.float32)
    s = SlaNetSession(SLANET, 0)
    loc, probs = s(x)
    rloc, rprobs = onnx_ref.run(SLANET, x)
    assert probs.shape == rprobs.shape and np.array_equal(probs.argmax(-1), rprobs.argmax(-1))
    assert np.abs(probs - rprobs).max() < 2e-4 and np.abs(loc - rloc).max() < 2e-4
    lengths = [int(np.argmax(r == 29)) for r in probs.argmax(-1)]
    assert len(set(lengths)) >= 4 and max(lengths) + 2 == probs.shape[1]


@pytest.mark.gpu
def test_slanet_loop_edges_max_steps_and_split_batches():
    """(1) a row that never emits the end token: the loop runs to max_steps and the outputs are the first max_steps rows of the
    free-running decode; (2) batches larger than MAX_BATCH are decoded in pieces, each stopping at its own step, rows padded
    with the untouched-row values."""
    from rapiddoc_b200.table import SlaNetSession
    _, x, _ = MG.table_inputs()
    s = SlaNetSession(SLANET, 0)
    loc, probs = s(x)                                   # 42 rows (longest table: 40 steps + eos + the untouched row)
    ids = probs.argmax(-1)
    s.max_steps, s.eos = 20, 3                          # token 3 ("</td>" of a spanning cell) never appears in these tables
    loc20, probs20 = s(x)
    assert s.last_steps == 20 and probs20.shape == (3, 20, 30) and loc20.shape == (3, 20, 4)
    assert np.array_equal(probs20.argmax(-1), ids[:, :20]) and np.abs(probs20 - probs[:, :20]).max() < 1e-6
    s.max_steps, s.eos = 501, 29
    s.MAX_BATCH = 2                                     # instance override: 3 tables -> pieces of 2 + 1
    loc_s, probs_s = s(x)
    assert probs_s.shape[0] == 3 and probs_s.shape[1] == 42          # piece 1 = tables 0, 1 -> 42 rows; piece 2 = table 2 -> 18, padded
    ids_s = probs_s.argmax(-1)
    for b in range(3):
        end = int(np.argmax(ids[b] == 29))
        assert np.array_equal(ids_s[b, :end + 1], ids[b, :end + 1]) and np.abs(loc_s[b, :end + 1] - loc[b, :end + 1]).max() < 1e-6
    assert np.allclose(probs_s[2, 18:], 1 / 30) and not loc_s[2, 18:].any()


@pytest.mark.gpu
def test_table_device_preprocess_is_bit_identical_to_the_host_class():
    import torch
    from rapiddoc_b200.table import TablePreprocess
    imgs, x, shapes = MG.table_inputs()
    pre = TablePreprocess()
    d, dshapes = pre.device_batch(imgs, 0)
    assert d.shape == (3, 488, 488, 4) and np.array_equal(dshapes, shapes)
    got = d.cpu().numpy()
    assert np.array_equal(got[..., :3].transpose(0, 3, 1, 2), x) and not got[..., 3].any()
    d2, _ = pre.device_batch(imgs[:1], 0)                                   # another batch size: its own staging canvases
    assert torch.equal(d2[0], d[0])


@pytest.mark.gpu
def test_table_structurer_interface():
    from rapiddoc_b200.table import B200TableStructurer, TableLabelDecode
    imgs, x, shapes = MG.table_inputs()
    ts = B200TableStructurer(device=0)
    structs, cells = ts(imgs)
    rloc, rprobs = onnx_ref.run(SLANET, x)
    dec = TableLabelDecode(onnx_lite.load(SLANET).meta["character"].splitlines(), slanet_plus=False, device=0)
    rstructs, rcells = dec(rloc, rprobs, shapes, imgs)
    for (tok, score), (rtok, rscore), cb, rcb, (rows, cols) in zip(structs, rstructs, cells, rcells, [(4, 3), (6, 4), (3, 2)]):
        assert tok == rtok and abs(score - rscore) < 1e-4
        assert tok.count("<tr>") == rows and tok.count("<td></td>") == rows * cols
        assert cb.shape == rcb.shape == (rows * cols, 4) and np.abs(cb - rcb).max() < 0.1        # pixels


# ------------------------------------------------------------- the OCR oracles against the reference's own ONNX model files
REF_RES = "/root/reference/rapid_doc/resources"


def test_ocr_oracles_match_the_reference_onnx_files():
    """RapidDoc's default engine is onnxruntime on ch_PP-OCRv6_{det,rec}_small.onnx; the repo's oracles are torch restatements on
    the safetensors weights.  An independent runtime (OpenCV DNN) on the det file and the node-by-node interpreter on the rec file
    agree with them."""
    if not os.path.isdir(REF_RES):
        pytest.skip("reference tree not mounted")
    cv2 = pytest.importorskip("cv2")
    import torch
    from oracle import nets, ocr_post as P
    from rapiddoc_b200 import synth
    page = synth.det_pages(1, 256, 512, seed=4, lines=5)[0]
    x = P.det_preprocess(page, limit_side_len=4096)
    want = nets.det_forward(x)
    net = cv2.dnn.readNetFromONNX(os.path.join(REF_RES, "ch_PP-OCRv6_det_small.onnx"))
    net.setInput(x)
    dnn = net.forward()
    assert np.abs(dnn - want).max() < 2e-4 and ((dnn > 0.3) != (want > 0.3)).sum() == 0
    assert np.abs(onnx_ref.run(os.path.join(REF_RES, "ch_PP-OCRv6_det_small.onnx"), x) - want).max() < 2e-5
    assert np.abs(GOLD["ocr_det_onnx_prob"].astype(np.float32) - want[0, 0]).max() < 1e-3            # fp16 storage
    xr = np.random.RandomState(0).randn(3, 3, 48, 160).astype(np.float32)
    probs = onnx_ref.run(os.path.join(REF_RES, "ch_PP-OCRv6_rec_small.onnx"), xr)
    mine = torch.softmax(torch.from_numpy(nets.rec_logits(xr)), -1).numpy()
    assert probs.shape == mine.shape == (3, 20, 18710)
    assert np.abs(probs - mine).max() < 2e-4 and np.array_equal(probs.argmax(-1), mine.argmax(-1))
    assert np.array_equal(GOLD["ocr_rec_onnx_ids"], mine.argmax(-1))


@pytest.mark.gpu
def test_cuda_engines_match_the_reference_onnx_goldens():
    """The CUDA det / rec engines (fp32 mode) against outputs of the reference's ONNX model files (fixtures made by cv2.dnn / the
    interpreter in the build container)."""
    from oracle import ocr_post as P
    from rapiddoc_b200 import PREC_FP32, synth
    from rapiddoc_b200.engine import DetEngine, RecEngine
    page = synth.det_pages(1, 256, 512, seed=4, lines=5)[0]
    det = DetEngine(device=0, precision=PREC_FP32)
    prob, bitmap = det.infer_u8(page[None])
    gold = GOLD["ocr_det_onnx_prob"].astype(np.float32)
    assert prob.shape[1:] == gold.shape and np.abs(prob[0] - gold).max() < 1e-3               # fp16 storage of the fixture
    far = np.abs(gold - 0.3) > 2e-3
    assert np.array_equal((prob[0] > 0.3)[far], (gold > 0.3)[far])
    xr = np.random.RandomState(0).randn(3, 3, 48, 160).astype(np.float32)
    out = RecEngine(device=0, precision=PREC_FP32).infer_f32(xr)
    assert np.array_equal(out["ids"], GOLD["ocr_rec_onnx_ids"])
    assert np.abs(out["probs"] - GOLD["ocr_rec_onnx_pmax"]).max() < 2e-4


@pytest.mark.gpu
def test_table_and_orientation_on_a_second_device():
    """Every op / the cluster decode kernel select their device themselves and put the caller's device back."""
    import torch
    from rapiddoc_b200 import _lib
    if _lib.load().rdb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from rapiddoc_b200.orientation import B200Orientation
    from rapiddoc_b200.table import B200TableStructurer
    imgs, x, shapes = MG.table_inputs()
    a, b = B200TableStructurer(device=0), B200TableStructurer(device=1)
    (sa, ca), (sb, cb) = a(imgs), b(imgs)
    assert [s[0] for s in sa] == [s[0] for s in sb] and all(np.array_equal(u, v) for u, v in zip(ca, cb))
    rots, _ = MG.orientation_inputs()
    assert np.array_equal(B200Orientation(device=1).scores(rots), B200Orientation(device=0).scores(rots))
    assert torch.cuda.current_device() == 0
