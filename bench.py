#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config, one process per GPU.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference ...                     (reference CPU arm, rank 0 only)

Default workload (config.workload) `ocr_pipeline_b64_1024x1024` = BASELINE.json's metric, "pages/sec (full det+rec
pipeline)", on configs[1]'s input (batch = 64 synthetic 1024x1024 pages per GPU): one STEP = one pass of the whole OCR
hot path over the batch — DBNet detection, DB post-process (boxes), text-line crops, LightSVTR recognition, CTC decode,
texts — through `B200OcrModel.ocr_pages` (the fused form of the reference's `_run_ocr_det_batch` +
`_run_ocr_rec_postprocess` window flow).  The line also carries `secondary` (the det-only and rec-only kernel-level
numbers of configs[1] / configs[2]), `parity` (decision flips against the CPU oracle on pages / crops of the timed
batch) and `fp32_exact` (the same pipeline at the reference's precision).
`--workload det` / `--workload rec` run configs[1] / configs[2] alone: det = fused normalise -> DBNet forward -> DB
binarise + 2x2 dilate over 64 pages; rec = 512 48x320 crops, forward + fused CTC greedy decode.

value = whole-job pages/s (crops/s) with the batch resident in HBM, timed with CUDA events
        on the stream the kernels are launched on, barrier + synchronize on both sides,
        max over ranks.
e2e   = the same metric through the C-ABI with HOST (pinned) buffers: H2D of the uint8
        pages and D2H of prob map + bitmap (det) / ids+probs+text (rec) inside the timed
        region.
roofline = the dominant kernel of the step (largest share of device time in a profiled
        pass taken right after the timed region, CUDA events around every launch on the
        launching stream): algorithmic bytes per launch / average launch duration against
        the measured HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline = the CPU oracle port of the same path (oracle/nets.py + oracle/ocr_post.py,
        torch CPU fp32, all host cores) on a bounded sample, rank 0 at N=1 only.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "pipeline": dict(name="ocr_pipeline_b64_1024x1024", batch=64, h=1024, w=1024, unit="pages/s", metric="pages/sec (full det+rec pipeline)"),
    "formula": dict(name="formula_ppformulanet_plus_m_b32_384", batch=32, h=384, w=384, unit="crops/s",
                    metric="formula crops/sec (PP-FormulaNet_plus-M: PPHGNetV2-B6 encoder + 64 greedy MBart tokens)"),
    "table": dict(name="table_slanet_1m_b32_488x488", batch=32, h=488, w=488, unit="tables/s",
                  metric="table crops/sec (SLANet-1m structure recognition: preprocess + backbone + GRU-attention decode + label decode)"),
    "det": dict(name="det_dbnet_b64_1024x1024", batch=64, h=1024, w=1024, unit="pages/s", metric="pages/sec (DBNet text-detection hot path)"),
    "rec": dict(name="rec_svtr_ctc_b512_48x320", batch=512, h=48, w=320, unit="crops/s", metric="SVTR text-line crops/sec"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("RDB_BENCH_WORKLOAD", "pipeline"), choices=list(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("RDB_BENCH_PRECISION", "fp16"), choices=["fp16", "fp32", "tf32"],
                    help="fp16 / fp32 for the OCR and formula workloads; the table workload computes in fp32 storage: fp32 (SIMT, default) or tf32 (tcgen05)")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rec-batch", type=int, default=int(os.environ.get("RDB_BENCH_REC_BATCH", "256")), help="Rec.rec_batch_num of the pipeline workload (both arms)")
    ap.add_argument("--no-stream", action="store_true", help="pipeline workload: issue the steps one by one (ocr_pages) instead of through ocr_pages_stream")
    ap.add_argument("--no-secondary", action="store_true", help="pipeline workload: skip the det-only / rec-only / fp32 legs")
    ap.add_argument("--chunk-pixels", type=int, default=0)
    ap.add_argument("--profile-out", default="", help="write the full per-kernel table of the profiled pass to this JSON file")
    return ap.parse_args()


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self._stop = index, [], None, None, False

    def start(self):
        """In-process NVML polling (one cheap query every 25 ms; a looping `nvidia-smi -lms` process costs the host-bound pipeline
        real CPU time and driver-lock contention); `nvidia-smi` is the fallback when pynvml cannot initialise."""
        if os.environ.get("RDB_BENCH_CLOCKS", "nvml") == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                phys = self.index
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                if vis and all(v.strip().isdigit() for v in vis.split(",")):
                    phys = int(vis.split(",")[self.index])
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
                self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
                pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
                self.nvml = pynvml
                self.t = threading.Thread(target=self._poll, daemon=True)
                self.t.start()
                return
            except Exception:
                self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        bits = [(0x8, 2), (0x40, 3), (0x20, 4), (0x4, 5)]        # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap (NVML reason masks)
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                row = [str(sm), str(self.max_sm), "Not Active", "Not Active", "Not Active", "Not Active"]
                for bit, col in bits:
                    if mask & bit:
                        row[col] = "Active"
                self.rows.append((time.time(), row))
            except Exception:
                pass
            time.sleep(0.025)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=5.0):
        """nvidia-smi needs ~100+ ms to start: block until the first sample is in."""
        t0 = time.time()
        while (self.proc is not None or self.nvml is not None) and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin=None, t_end=None):
        """Summarise the samples taken inside [t_begin, t_end] (the timed region)."""
        if self.proc is None and self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        if self.nvml is not None:
            self._stop = True
            self.t.join(timeout=1.0)
        else:
            self.proc.terminate()
        inside = [r for t, r in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.03)]
        self.rows = inside if inside else [r for _, r in self.rows[-3:]]
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------- roofline model
def _kv(name):
    m = re.match(r"([^\[]+)(?:\[(.*)\])?", name)
    kind, args = m.group(1), {}
    if m.group(2):
        for item in m.group(2).split(","):
            k, v = item.split("=")
            args[k] = int(v)
    return kind, args


def algorithmic_bytes(name, esz, wl, n_chunk):
    """Minimum HBM bytes one launch of this kernel must move: every input
    element read once, every output written once, at the storage size of the precision mode."""
    kind, a = _kv(name)
    H, W = wl["h"], wl["w"]
    base = kind.replace("_tcp", "").replace("_tc", "").replace("_tiled", "")
    if base.startswith("gemm") and "ctc" in base:
        return a["M"] * a["K"] * esz + a["N"] * a["K"] * 2 + a["M"] * 12 * ((a["N"] + 255) // 256) * 4
    if base.startswith("gemm"):
        return a["M"] * (a["K"] + a["N"] + (a["N"] if a.get("res") else 0)) * esz + a["N"] * a["K"] * esz
    if base.startswith("dwconv"):
        return a["P"] * a["C"] * esz * (1 + a.get("s", 1))       # stride-s input is s x the output
    if base in ("stem2a", "stem2b", "head_conv3x3") and "P" in a:
        return a["P"] * (a["C"] + a["N"]) * esz                    # stride-1 dense convs: in + out once
    if base in ("stem_fused", "stem_planar"):
        c1 = 24 if wl["h"] > 48 else 48
        return a["P"] * (4 * c1 + 2 * c1) * esz                    # e1 read once (4 px per output px), stem4 output written once
    if base == "stem1" and "P" in a:
        return a["P"] * (4 * 3 + a["N"] * esz)                     # uint8 page (4 input px per output px) in, C1-channel map out
    if base == "mlp":
        return a["M"] * (a["C"] + a["N"] + (a["N"] if a.get("res") else 0)) * esz    # block input (+ residual re-read) and output, once each
    if base == "head_planar":
        return int(a["P"] * (24 * esz * (1 + 1 / 4 + 1 / 16 + 1 / 64) + 16 * 5))   # the four 24-ch pyramid maps read once, 4x4 prob f32 + seg u8 written per pixel
    if base == "stem3" and "P" in a:
        return (4 * a["P"] * a["C"] + a["P"] * a["N"]) * esz       # 3x3 stride 2
    if base == "head_tail" and "P" in a:
        return a["P"] * 24 * esz + a["P"] * 16 * 5                 # 24-ch map in, 4x4 prob f32 + seg u8 out per pixel
    if base == "se_pool":
        return a["P"] * a["C"] * esz
    if base == "se_scale":
        return 2 * a["P"] * a["C"] * esz
    px = n_chunk * H * W
    table = {
        "stem1": px * 3 * 1 + px // 4 * (24 if wl["h"] > 48 else 48) * esz, "stem2a": px // 4 * (24 + 12) * esz, "stem2b": px // 4 * (12 + 24) * esz,
        "stem_pool": px // 4 * 48 * esz, "stem3": px // 4 * 48 * esz + px // 16 * 24 * esz,
        "head_conv3x3": px // 16 * (96 + 24) * esz, "head_tail": px // 16 * 24 * esz + px * 5,
        "db_dilate": px * 2, "neck_concat": px // 16 * 96 * esz * 2,
    }
    return table.get(base)


def algorithmic_flops(name):
    kind, a = _kv(name)
    if kind.startswith("gemm"):
        return 2 * a["M"] * a["K"] * a["N"]
    return None


# ------------------------------------------------------------------------------- CPU oracle arm
def cpu_step_det(pages):
    from oracle import nets, ocr_post as P
    x = np.concatenate([P.det_preprocess(p, limit_side_len=max(p.shape[:2])) for p in pages])
    prob = nets.det_forward(x)
    return [P.db_bitmap(prob[i, 0], 0.3, True) for i in range(len(pages))]


def cpu_step_rec(crops):
    from oracle import nets, ocr_post as P
    x, _ = P.rec_batch_tensor(list(crops))
    probs = nets.rec_forward(x)
    return P.ctc_decode(probs, nets.load_characters())


def pick_cpu_threads(wl_key, data):
    """The oracle runs on torch CPU.  "All host cores" is what the reference does
    (torch default intra-op threads), but on many-core hosts that oversubscribes the small
    convolutions, so the baseline uses the FASTEST of {all, 64, 32, 16} threads on a tiny probe —
    the most favourable setting for the CPU side."""
    import torch
    fn = cpu_step_det if wl_key == "det" else cpu_step_rec
    probe = data[:1] if wl_key == "det" else data[:16]
    best, best_t = None, None
    for th in sorted({os.cpu_count(), 64, 32, 16}, reverse=True):
        if th > os.cpu_count():
            continue
        torch.set_num_threads(th)
        fn(probe)
        t0 = time.perf_counter()
        fn(probe)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = th, dt
    torch.set_num_threads(best)
    return best


def time_cpu(wl_key, data, sample, reps=1):
    fn = cpu_step_det if wl_key == "det" else cpu_step_rec
    fn(data[: min(2, sample)])  # warm
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(data[:sample])
    dt = (time.perf_counter() - t0) / reps
    return sample / dt, dt


# ------------------------------------------------------------------------------- pipeline workload
PIPE = dict(limit_side_len=1024, box_thresh=0.3, unclip_ratio=1.8, merge=True)   # page OCR settings of model_init.py:106-111


def cpu_step_pipeline(pages, rec_batch, detail=False):
    from oracle import pipeline as OP
    return OP.ocr_pages(list(pages), limit_side_len=PIPE["limit_side_len"], box_thresh=PIPE["box_thresh"], unclip_ratio=PIPE["unclip_ratio"],
                        merge=PIPE["merge"], rec_batch_num=rec_batch, return_detail=detail)


def make_pipeline_model(device, prec, rec_batch, blobs=None):
    from rapiddoc_b200.ocr import B200OcrModel
    return B200OcrModel(det_db_box_thresh=PIPE["box_thresh"], det_db_unclip_ratio=PIPE["unclip_ratio"], enable_merge_det_boxes=PIPE["merge"],
                        ocr_config={"Det.limit_side_len": PIPE["limit_side_len"], "Rec.rec_batch_num": rec_batch}, device=device,
                        precision=prec, blobs=blobs)


def parity_report(model, pages, want, detail):
    """Decision-level differences between the B200 pipeline and the CPU oracle on the same pages (flips are
    counted and reported, not hidden in a tolerance)."""
    n = len(pages)
    rep = {"pages_checked": n}
    # detection maps
    prob, bitmap = model.text_detector.engine.infer_u8(np.stack(pages), thresh=0.3, use_dilation=True)
    oprob = np.stack([m[0] for m in detail["maps"]])
    obm = np.stack([m[1] for m in detail["maps"]])
    rep["max_abs_dprob"] = float(np.abs(prob - oprob).max())
    rep["bitmap_flips"] = int((bitmap != obm).sum())
    rep["bitmap_pixels"] = int(obm.size)
    # boxes (after sort / merge), end-to-end texts
    got = model.ocr_pages(list(pages))
    box_mismatch, box_count_diff, box_max_px, text_mismatch, lines = 0, 0, 0.0, 0, 0
    for g, w in zip(got, want):
        g, w = g or [], w or []
        box_count_diff += abs(len(g) - len(w))
        for (gb, (gt, _)), (wb, (wt, _)) in zip(g, w):
            d = float(np.abs(np.asarray(gb, np.float64) - np.asarray(wb, np.float64)).max())
            box_mismatch += d > 0
            box_max_px = max(box_max_px, d)
            text_mismatch += gt != wt
            lines += 1
    rep.update(lines_checked=lines, box_count_diff=box_count_diff, box_mismatch=int(box_mismatch), box_max_px=box_max_px,
               text_mismatch=int(text_mismatch))
    # recognition on the ORACLE's crops (identical inputs for both sides): per-step argmax flips, decoded text, confidence
    crops = detail["crops"]
    if crops:
        model.text_recognizer.keep_ids = True
        r = model.text_recognizer(crops)
        model.text_recognizer.keep_ids = False
        flips = sum(int((np.asarray(a) != np.asarray(b)).sum()) for a, b in zip(model.text_recognizer.last_ids, detail["ids"]))
        steps = sum(len(b) for b in detail["ids"])
        rep.update(crops_checked=len(crops), ctc_steps_checked=int(steps), argmax_flips=int(flips),
                   crop_text_mismatch=int(sum(a != b[0] for a, b in zip(r.txts, detail["rec"]))),
                   max_abs_dconf=float(max(abs(a - b[1]) for a, b in zip(r.scores, detail["rec"]))))
    return rep


def aggregate_roofline(prof, esz, wl_det, wl_rec, hbm, peak_src):
    """The pipeline step launches each kernel at many shapes (every recognition batch has its own width): kernels are grouped
    by family, achieved = sum(algorithmic bytes) / sum(device time) over all launches of the dominant family."""
    fam = {}
    total = sum(v[0] for v in prof.values()) or 1.0
    for name, (ms, cnt) in prof.items():
        kind, a = _kv(name)
        wl = wl_det if (a.get("P", 0) >= 1 << 18 or kind.startswith(("head_planar", "stem_planar", "db_", "dwconv7", "se_", "neck"))) else wl_rec
        ab = algorithmic_bytes(name, esz, wl, 16)
        f = fam.setdefault(kind, {"ms": 0.0, "launches": 0, "bytes": 0.0, "modelled_ms": 0.0})
        f["ms"] += ms
        f["launches"] += cnt
        if ab:
            f["bytes"] += float(ab) * cnt
            f["modelled_ms"] += ms
    modelled = {k: v for k, v in fam.items() if v["bytes"] > 0}
    kind, f = max((modelled or fam).items(), key=lambda kv: kv[1]["ms"])
    ach = f["bytes"] / (f["modelled_ms"] / 1e3) / 1e9 if f["modelled_ms"] else None
    traffic = traffic_src = None
    try:        # DRAM bytes per launch of this family from the committed ncu --set full capture (recogniser chunk of the same step)
        for row in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))):
            if row.get("kernel") == "family:" + kind:
                traffic, traffic_src = row["dram_bytes"], row["source"]
    except Exception:
        pass
    return {"kernel": kind, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": (ach / hbm) if ach else None, "traffic": traffic,
            "traffic_source": traffic_src, "peak_source": peak_src, "launches_profiled": f["launches"], "avg_launch_us": f["ms"] / f["launches"] * 1e3,
            "algorithmic_bytes_per_launch": f["bytes"] / f["launches"], "share_of_step": f["ms"] / total,
            "top5": [{"kernel": k, "share": v["ms"] / total, "launches": v["launches"], "avg_us": v["ms"] / v["launches"] * 1e3,
                      "gbps": (v["bytes"] / (v["modelled_ms"] / 1e3) / 1e9) if v["modelled_ms"] else None}
                     for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])[:5]]}


def run_pipeline(args, wl):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from rapiddoc_b200 import _lib, PREC_FP16, PREC_FP32, synth, weights as W
    from rapiddoc_b200.parallel import broadcast_blob, pin_rank_to_cores
    cores = pin_rank_to_cores(local, world) if world > 1 else None
    prec = PREC_FP16 if args.precision == "fp16" else PREC_FP32
    esz = 2 if prec == PREC_FP16 else 4
    blobs = (broadcast_blob(W.det_blob() if rank == 0 else None, local), broadcast_blob(W.rec_blob() if rank == 0 else None, local))
    B, H, Wd = wl["batch"], wl["h"], wl["w"]
    uniq = min(B, 8)
    base = synth.det_pages(uniq, H, Wd, seed=1 + rank)
    host = torch.empty((B, H, Wd, 3), dtype=torch.uint8).pin_memory()
    for i in range(B):   # distinct pages: the unique ones rolled by a per-page offset (text lines move, content stays text)
        host[i] = torch.from_numpy(np.roll(base[i % uniq], shift=(7 * (i // uniq), 13 * (i // uniq)), axis=(0, 1)))
    host_np = host.numpy()
    pages = [host_np[i] for i in range(B)]
    dev_pages = host.cuda()
    model = make_pipeline_model(local, prec, args.rec_batch, blobs)
    det_t, rec_t = model.text_detector, model.text_recognizer

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reset_stats():
        for t in (det_t, rec_t):
            for k in t.stats:
                t.stats[k] = 0

    # ---- value: pages resident in HBM, CUDA events on the launching (torch current) stream
    for _ in range(max(args.warmup, 3)):
        res = model.ocr_pages(dev_pages)
    if not args.no_stream:      # warm the timed path itself too: the streaming stage threads, their pinned staging and streams
        for _ in model.ocr_pages_stream(dev_pages for _ in range(max(args.warmup, 3))):
            pass
    lines_per_step = sum(len(r or []) for r in res)
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    barrier()
    reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    if args.no_stream:
        for _ in range(args.steps):
            model.ocr_pages(dev_pages)
    else:       # steps issued through the streaming API: detection of step k+1 overlaps recognition of step k
        for _ in model.ocr_pages_stream(dev_pages for _ in range(args.steps)):
            pass
    e1.record()
    barrier()
    clocks = sampler.stop(t_begin, time.time())
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = det_t.stats["launches"] + rec_t.stats["launches"]
    crops_per_step = rec_t.stats["crops"] / args.steps
    value = world * B * args.steps / (ms / 1e3)
    # ---- e2e: host (pinned) pages in, python results out; every copy inside the timed region
    for _ in range(2):
        model.ocr_pages(pages)
    if not args.no_stream:
        for _ in model.ocr_pages_stream(pages for _ in range(2)):
            pass
    barrier()
    reset_stats()
    t0 = time.perf_counter()
    if args.no_stream:
        for _ in range(args.steps):
            model.ocr_pages(pages)
    else:
        for _ in model.ocr_pages_stream(pages for _ in range(args.steps)):
            pass
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = world * B * args.steps / dt
    h2d = (det_t.stats["h2d_bytes"] + rec_t.stats["h2d_bytes"]) / args.steps
    d2h = (det_t.stats["d2h_bytes"] + rec_t.stats["d2h_bytes"]) / args.steps
    line = None
    if rank == 0:
        # ---- roofline: per-kernel device time of one profiled step (events around every launch), grouped by kernel family
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        os.environ["RDB_LANES"] = "1"
        _lib.load().rdb_switches_reload()
        _lib.profile(True)
        _lib.profile_reset()
        model.ocr_pages(dev_pages)
        torch.cuda.synchronize()
        prof = _lib.profile_dump()
        _lib.profile(False)
        os.environ.pop("RDB_LANES", None)
        _lib.load().rdb_switches_reload()
        roofline = aggregate_roofline(prof, esz, WORKLOADS["det"], WORKLOADS["rec"], hbm,
                                      "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)")
        if args.profile_out:
            tot = sum(v[0] for v in prof.values())
            json.dump({"workload": wl["name"], "precision": args.precision, "total_ms": tot,
                       "kernels": [{"kernel": k, "total_ms": v[0], "launches": v[1], "share": v[0] / tot} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])]},
                      open(args.profile_out, "w"), indent=1)
        cpu = parity = secondary = fp32 = None
        if world == 1 and not args.no_cpu_baseline:
            # ---- CPU baseline + parity: the oracle port of the same window flow on a bounded sample of this step's pages
            sample = 4
            threads = pick_cpu_threads("det", host_np[:1])
            t0 = time.perf_counter()
            want, detail = cpu_step_pipeline(pages[:sample], args.rec_batch, detail=True)
            dt_cpu = time.perf_counter() - t0
            cpu = {"value": sample / dt_cpu, "unit": wl["unit"], "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
                   "sample": f"{sample} pages of this step's batch through the CPU oracle port of the whole flow (torch CPU fp32 nets + OpenCV + "
                             f"CTC decode, rec_batch_num {args.rec_batch}), {dt_cpu:.1f}s"}
            parity = parity_report(model, pages[:sample], want, detail)
            if not args.no_secondary and prec == PREC_FP16:
                m32 = make_pipeline_model(local, PREC_FP32, args.rec_batch, blobs)
                m32.ocr_pages(dev_pages)                 # warm-up at the timed shape (workspace pools, staging buffers)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(2):
                    m32.ocr_pages(dev_pages)
                torch.cuda.synchronize()
                fp32 = {"value": 2 * B / (time.perf_counter() - t0), "unit": wl["unit"], "note": "same step in RDB_PREC_FP32 (fp32 storage + fp32 SIMT math, the reference's precision), 2 steps, step by step",
                        "parity": parity_report(m32, pages[:sample], want, detail)}
                del m32
        if not args.no_secondary:
            secondary = secondary_numbers(model, dev_pages, B, H, Wd)
        line = {"metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16" if prec == PREC_FP16 else "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "batch_per_gpu": B, "h": H, "w": Wd, "precision": args.precision, "input": "uint8 BGR HWC pages",
                           "flow": "det (limit_side_len 1024, thresh .3, box_thresh .3, unclip 1.8) -> sorted/merged boxes -> get_rotate_crop_image -> rec",
                           "rec_batch_num": args.rec_batch, "text_lines_per_step": int(lines_per_step),
                           "issue": "step by step (ocr_pages)" if args.no_stream else "streaming API (ocr_pages_stream): detection stage of step k+1 overlaps recognition of step k; every step's results are materialised on the host inside the timed region", "crops_per_step": crops_per_step,
                           "l2": f"inputs {host.numel() / 1e6:.0f} MB + activations per step exceed the 126 MB L2 (no explicit flush)",
                           "parallelism": f"page-parallel replicas x{world}, NCCL weight broadcast at init only",
                           "host_cores_per_rank": len(cores) if cores else os.cpu_count()},
                "clocks": clocks, "e2e": {"value": e2e, "unit": wl["unit"], "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "fp32_exact": fp32, "secondary": secondary,
                "skipped": SKIPPED_CONFIGS}
        emit_line(json.dumps(line), flush=True)
    if world > 1:
        barrier()
        dist.destroy_process_group()


SKIPPED_CONFIGS = {
    "configs[0] single 960x960 page PP-DocLayout-S via onnxruntime CPU": "skipped: weights unavailable (PP-DocLayout-S is downloaded at first run; not on disk) and onnxruntime is not installed",
    "configs[3] SLANet_plus + UNET table structure, batch=32 488x488": "skipped: weights unavailable (SLANet_plus / UNET model files are downloaded at first run; not on disk) — the same shape runs on the SLANet RapidDoc ships (slanet-1m.onnx, ModelType.SLANET1M) as `bench.py --workload table`",
    "configs[4] full PP-DocLayoutV3 + OCRv5 det/rec + FormulaNet_plus-M, 1200 A4 pages": "skipped: weights unavailable for layout / formula / table (only the OCR det+rec part runs: this line)",
}


def secondary_numbers(model, dev_pages, B, H, Wd, steps=5):
    """det-only (configs[1]) and rec-only (configs[2]) device-resident numbers of the same engines, a few steps each."""
    import torch
    from rapiddoc_b200 import synth
    det, rec = model.text_detector.engine, model.text_recognizer.engine
    d_prob = torch.empty((B, H, Wd), dtype=torch.float32, device="cuda")
    d_bm = torch.empty((B, H, Wd), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()

    def timed(fn, units):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return units * steps / (e0.elapsed_time(e1) / 1e3)
    out = {"det_pages_per_s": timed(lambda: det.infer_u8(dev_pages, prob=d_prob, bitmap=d_bm, stream=st), B), "det_config": f"{B} pages {H}x{Wd}, forward + DB bitmap, device-resident"}
    del d_prob, d_bm
    n = 512
    crops = torch.from_numpy(synth.rec_crops(n, 48, 320, seed=2)).cuda()
    vw = torch.full((n,), 320, dtype=torch.int32, device="cuda")
    outs = rec._outs(n, rec.tokens(320), crops, False)
    out["rec_crops_per_s"] = timed(lambda: rec.infer_u8(crops, vw, stream=st, outs=outs), n)
    out["rec_config"] = f"{n} crops 48x320, forward + fused CTC decode, device-resident"
    return out


# ------------------------------------------------------------------------------- formula workload
FORMULA_TOKENS = 64
ENC_GFLOP_PER_CROP = 98.94          # BASELINE.md section 2 (FlopCounterMode on the reference module, 1x1x384x384)


def formula_inputs(n, seed=4):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 1, 384, 384), np.float32)
    for i in range(n):
        lvl, amp = rng.uniform(-3.0, 1.19), rng.uniform(0.2, 2.0)
        x[i, 0] = (lvl + amp * np.kron(rng.standard_normal((12, 12)), np.ones((32, 32))) + 0.3 * rng.standard_normal((384, 384))).astype(np.float32)
    return x


def run_formula(args, wl):
    """F3/F4 workload: 32 synthetic 384x384 formula crops per step through FormulaEngine (seeded synthetic weights of the exact
    PP-FormulaNet_plus-M architecture: the trained checkpoint is not available offline), encoder + 64 greedy tokens."""
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from rapiddoc_b200 import PREC_FP16, PREC_FP32, formula as FM
    prec = PREC_FP16 if args.precision == "fp16" else PREC_FP32
    sd = FM.synthetic_state_dict(seed=0)
    eng = FM.FormulaEngine(sd, device=local, precision=prec, max_new_tokens=FORMULA_TOKENS, sync_every=64)
    B = wl["batch"]
    x_host = torch.from_numpy(formula_inputs(B, seed=4 + rank)).pin_memory()
    x_dev = x_host.cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def mx(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    for _ in range(max(args.warmup, 3)):
        ids = eng(x_dev)
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    barrier()
    eng.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        eng(x_dev)
    e1.record()
    barrier()
    clocks = sampler.stop(t_begin, time.time())
    ms = mx(e0.elapsed_time(e1))
    launches = eng.launches
    value = world * B * args.steps / (ms / 1e3)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids = eng(x_host.numpy())          # host float32 crops in, ids out
    torch.cuda.synchronize()
    dt = mx(time.perf_counter() - t0)
    e2e = world * B * args.steps / dt
    if rank == 0:
        # encoder alone: the dense-contraction part (tensor roofline)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(3):
            enc = eng.encode(x_dev)
        a1.record()
        torch.cuda.synchronize()
        enc_ms = a0.elapsed_time(a1) / 3
        if args.profile_out:
            from rapiddoc_b200 import _lib
            _lib.profile(True)
            _lib.profile_reset()
            eng.encode(x_dev)
            torch.cuda.synchronize()
            prof = _lib.profile_dump()           # synchronises and resolves the pending event pairs
            _lib.profile(False)
            tot = sum(v[0] for v in prof.values()) or 1.0
            json.dump({"workload": wl["name"], "precision": args.precision, "encoder_total_ms": tot,
                       "kernels": [{"kernel": k, "total_ms": v[0], "launches": v[1], "share": v[0] / tot} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])]},
                      open(args.profile_out, "w"), indent=1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0) if prec == PREC_FP16 else None
        ach = ENC_GFLOP_PER_CROP * B / enc_ms        # GFLOP / ms = TFLOP/s
        roofline = {"kernel": "PPHGNetV2-B6 encoder (im2col + gemm_tc, all layers)" if prec == PREC_FP16 else "encoder (fp32 SIMT GEMM)", "bound": "tensor",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if peak else None, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured; fp16 and bf16 share the tcgen05 rate)" if peaks else "fallback",
                    "encoder_ms_per_step": enc_ms, "decoder_ms_per_step": ms / args.steps - enc_ms, "algorithmic_gflop_per_crop": ENC_GFLOP_PER_CROP}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import formula_net as FN
            import torch as _t
            _t.set_num_threads(min(os.cpu_count(), 32))
            sample = 2
            t0 = time.perf_counter()
            want, _ = FN.forward(x_host.numpy()[:sample], sd, FM.ARCH_M, FORMULA_TOKENS)
            dt_cpu = time.perf_counter() - t0
            n = min(want.shape[1], ids.shape[1])
            cpu = {"value": sample / dt_cpu, "unit": wl["unit"], "cores": min(os.cpu_count(), 32), "host_cores": os.cpu_count(), "kind": "port",
                   "sample": f"{sample} crops of this step's batch through the CPU oracle (torch fp32, BN unfolded, {FORMULA_TOKENS} greedy tokens), {dt_cpu:.1f}s",
                   "token_mismatch_vs_oracle": int((want[:, :n] != ids[:sample, :n]).sum()), "tokens_compared": int(want[:, :n].size)}
        line = {"metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16" if prec == PREC_FP16 else "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "batch_per_gpu": B, "h": 384, "w": 384, "precision": args.precision, "greedy_tokens": FORMULA_TOKENS,
                           "weights": "seeded synthetic weights of the exact architecture (trained checkpoint unavailable offline)",
                           "l2": "activations per step (GBs) exceed the 126 MB L2", "parallelism": f"crop-parallel replicas x{world}"},
                "clocks": clocks, "e2e": {"value": e2e, "unit": wl["unit"], "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": int(ids.size * 8)},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        emit_line(json.dumps(line), flush=True)
    if world > 1:
        barrier()
        dist.destroy_process_group()


def table_inputs(n, seed=0):
    from rapiddoc_b200 import synth
    return [synth.table_image(seed * 1000 + i, 3 + i % 6, 2 + i % 4, 300 + 10 * (i % 5), 400 + 16 * (i % 7), lines=(i % 3 != 0)) for i in range(n)]


def run_table(args, wl):
    """T3-T5 workload (BASELINE configs[3] shape: batch 32 of 488x488 table crops) on the weights RapidDoc ships (slanet-1m.onnx;
    SLANet_plus / UNET are downloaded at first run and are not on disk).  value: preprocessed crops resident in HBM -> (boxes,
    structure probabilities) on the device + the stop step read back; e2e: BGR crops on the host -> html tokens + cell boxes."""
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from rapiddoc_b200 import _lib, table as TB
    tprec = "tf32" if args.precision == "tf32" else "fp32"
    ts = TB.B200TableStructurer(device=local, precision=tprec)
    B = wl["batch"]
    imgs = table_inputs(B, seed=rank)
    x, shapes = ts.preprocess_op(imgs)
    x_host = torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.float32))).pin_memory()
    x_dev = x_host.cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def mx(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    for _ in range(max(args.warmup, 3)):
        loc, probs = ts.session(x_dev)
    steps_decoded = ts.session.last_steps
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    barrier()
    ts.session.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        ts.session(x_dev)
    e1.record()
    barrier()
    clocks = sampler.stop(t_begin, time.time())
    ms = mx(e0.elapsed_time(e1))
    launches = ts.session.launches
    value = world * B * args.steps / (ms / 1e3)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        structs, cells = ts(imgs)                 # host BGR crops in, html tokens + cell boxes out
    torch.cuda.synchronize()
    dt = mx(time.perf_counter() - t0)
    e2e = world * B * args.steps / dt
    if rank == 0:
        _lib.profile(True)
        _lib.profile_reset()
        ts.session(x_dev)
        torch.cuda.synchronize()
        prof = _lib.profile_dump()
        _lib.profile(False)
        tot = sum(v[0] for v in prof.values()) or 1.0
        fam = {}
        for k, v in prof.items():
            f = k.split("[")[0]
            fam.setdefault(f, [0.0, 0])
            fam[f][0] += v[0]
            fam[f][1] += v[1]
        top = sorted(fam.items(), key=lambda kv: -kv[1][0])
        if args.profile_out:
            json.dump({"workload": wl["name"], "total_ms": tot, "kernels": [{"kernel": k, "total_ms": v[0], "launches": v[1], "share": v[0] / tot}
                                                                             for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])]}, open(args.profile_out, "w"), indent=1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = peaks.get("hbm_gbs_sustained", peaks.get("hbm_gbs", 6500.0))
        # the backbone's convolutions dominate: fp32 activations; a GEMM launch must read A [M,K] and W [N,K] and write [M,N] once
        conv_ms = sum(v[0] for k, v in fam.items() if k in ("gemm_simt_op", "gemm_tf32_op", "im2col", "im2col_op", "dwconv_op", "chain_op"))
        g_bytes = g_ms = 0.0
        g_launches = 0
        for k, v in prof.items():
            kind, a = _kv(k)
            if kind in ("gemm_simt_op", "gemm_tf32_op"):
                g_bytes += 4.0 * (a["M"] * a["K"] + a["M"] * a["N"] + a["N"] * a["K"]) * v[1]
                g_ms += v[0]
                g_launches += v[1]
        ach = g_bytes / (g_ms * 1e-3) / 1e9 if g_ms else None
        roofline = {"kernel": ("gemm_tf32_op (tcgen05 kind::tf32 GEMM on fp32 storage" if tprec == "tf32" else "gemm_simt_op (fp32 SIMT GEMM") + ": the backbone's pointwise / im2col convolutions)", "bound": "hbm", "achieved": ach, "peak": hbm,
                    "unit": "GB/s", "frac": (ach / hbm) if ach else None, "traffic": None, "launches_profiled": g_launches,
                    "avg_launch_us": (g_ms * 1e3 / g_launches) if g_launches else None,
                    "algorithmic_bytes_per_launch": (g_bytes / g_launches) if g_launches else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback",
                    "share_of_step": (g_ms / tot) if tot else None, "top5": [{"kernel": k, "ms": v[0], "launches": v[1], "share": v[0] / tot} for k, v in top[:5]],
                    "backbone_ms": conv_ms, "profiled_total_ms": tot, "decode_steps": steps_decoded,
                    "note": "fp32 SIMT path (the reference's precision); the decode loop is latency-bound (one CTA per table, sequential steps)"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import onnx_ref
            import torch as _t
            _t.set_num_threads(min(os.cpu_count(), 32))
            sample = 2
            t0 = time.perf_counter()
            rloc, rprobs = onnx_ref.run(ts.session.net.path, x_host.numpy()[:sample])
            dt_cpu = time.perf_counter() - t0
            got_loc, got_probs = ts.session(x_dev[:sample])
            n = min(rprobs.shape[1], got_probs.shape[1])
            cpu = {"value": sample / dt_cpu, "unit": wl["unit"], "cores": min(os.cpu_count(), 32), "host_cores": os.cpu_count(), "kind": "port",
                   "sample": f"{sample} tables of this step's batch through the CPU oracle (node-by-node torch fp32 execution of the ONNX file incl. its Loop), {dt_cpu:.1f}s",
                   "token_mismatch_vs_oracle": int((rprobs[:, :n].argmax(-1) != got_probs[:, :n].argmax(-1)).sum()) + abs(rprobs.shape[1] - got_probs.shape[1]),
                   "tokens_compared": int(rprobs[:, :n].shape[0] * n), "max_abs_dprob": float(np.abs(rprobs[:, :n] - got_probs[:, :n]).max()),
                   "max_abs_dbox": float(np.abs(rloc[:, :n] - got_loc[:, :n]).max())}
        line = {"metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if tprec == "tf32" else "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "batch_per_gpu": B, "h": 488, "w": 488, "precision": tprec, "decode_steps": steps_decoded,
                           "weights": "slanet-1m.onnx (shipped with RapidDoc)", "tables": "synthetic ruled / borderless grids, 3-8 rows x 2-5 columns",
                           "l2": "activations per step exceed the 126 MB L2", "parallelism": f"crop-parallel replicas x{world}"},
                "clocks": clocks, "e2e": {"value": e2e, "unit": wl["unit"], "h2d_bytes_per_step": int(x_host.numel() * 4),
                                          "d2h_bytes_per_step": int(loc.size * 4 + probs.size * 4)},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        emit_line(json.dumps(line), flush=True)
    if world > 1:
        barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------- main arms
def run_reference(args, wl_key, wl):
    """Reference arm: the reference's own CPU implementation of the path.  RapidDoc is pure
    Python and cannot travel to the GPU box (its engines onnxruntime/rapidocr are absent even
    here), so this times the CPU oracle port (bit-identical to the reference's torch nets,
    tests/test_oracle.py) on all host cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rapiddoc_b200 import synth
    if wl_key == "formula":
        from oracle import formula_net as FN
        from rapiddoc_b200 import formula as FM
        import torch
        torch.set_num_threads(min(os.cpu_count(), 32))
        sd = FM.synthetic_state_dict(seed=0)
        x = formula_inputs(2)
        for _ in range(min(args.warmup, 1)):
            FN.forward(x[:1], sd, FM.ARCH_M, 4)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            FN.forward(x, sd, FM.ARCH_M, FORMULA_TOKENS)
        dt = time.perf_counter() - t0
        v = 2 * args.steps / dt
        emit_line(json.dumps({"impl": "reference", "metric": wl["metric"], "value": v, "unit": wl["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": wl["name"], "sample": "2 of the 32 per step", "engine": "torch CPU fp32 (oracle port of the reference module)"},
                          "cpu_baseline": {"value": v, "unit": wl["unit"], "cores": min(os.cpu_count(), 32), "host_cores": os.cpu_count(), "kind": "port", "sample": f"2 crops per step x {args.steps} steps"},
                          "e2e": {"value": v, "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    if wl_key == "table":
        from oracle import onnx_ref
        from rapiddoc_b200 import table as TB
        import torch
        torch.set_num_threads(min(os.cpu_count(), 32))
        path = os.path.join(ROOT, "weights", "slanet-1m.onnx")
        imgs = table_inputs(2)
        pre = TB.TablePreprocess()
        chars = onnx_ref.onnx_lite.load(path).meta["character"].splitlines()

        def step():
            x, shapes = pre(imgs)
            loc, probs = onnx_ref.run(path, np.asarray(x, np.float32))
            return loc, probs
        for _ in range(min(args.warmup, 1)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
        v = 2 * args.steps / dt
        emit_line(json.dumps({"impl": "reference", "metric": wl["metric"], "value": v, "unit": wl["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": wl["name"], "sample": "2 of the 32 per step", "engine": "torch CPU fp32 node-by-node execution of slanet-1m.onnx (onnxruntime is not installed)"},
                          "cpu_baseline": {"value": v, "unit": wl["unit"], "cores": min(os.cpu_count(), 32), "host_cores": os.cpu_count(), "kind": "port", "sample": f"2 tables per step x {args.steps} steps"},
                          "e2e": {"value": v, "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    sample = {"det": 4, "rec": 64, "pipeline": 2}[wl_key]
    data = synth.rec_crops(sample, wl["h"], wl["w"], seed=2) if wl_key == "rec" else synth.det_pages(sample, wl["h"], wl["w"], seed=1)
    fn = {"det": cpu_step_det, "rec": cpu_step_rec, "pipeline": lambda d: cpu_step_pipeline(d, args.rec_batch)}[wl_key]
    threads = pick_cpu_threads("rec" if wl_key == "rec" else "det", data)
    for _ in range(args.warmup):
        fn(data)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(data)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": wl["metric"], "value": v, "unit": wl["unit"], "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": f"{sample} of the {wl['batch']} per step", "engine": "torch CPU fp32 (oracle port of the reference nets)",
                       **({"rec_batch_num": args.rec_batch} if wl_key == "pipeline" else {})},
            "cpu_baseline": {"value": v, "unit": wl["unit"], "cores": threads, "host_cores": os.cpu_count(), "kind": "port", "sample": f"{sample} {wl['unit'].split('/')[0]} per step x {args.steps} steps"},
            "e2e": {"value": v, "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(json.dumps(line), flush=True)


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its version banner to fd 1 from C).
    Keep a private duplicate of the real stdout for the result line and point fd 1 (and Python's sys.stdout) at stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit_line(text, flush=True):
    out = _REAL_STDOUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    args = parse()
    claim_stdout()
    # NCCL prints its version banner (levels VERSION and WARN) and every debug line to stdout: route them to stderr so that
    # stdout carries the one JSON line only
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    wl_key = args.workload
    wl = dict(WORKLOADS[wl_key])
    if args.batch:
        wl["batch"] = args.batch
    if args.impl == "reference":
        return run_reference(args, wl_key, wl)
    if wl_key == "pipeline":
        return run_pipeline(args, wl)
    if wl_key == "formula":
        return run_formula(args, wl)
    if wl_key == "table":
        return run_table(args, wl)

    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from rapiddoc_b200 import _lib, PREC_FP16, PREC_FP32, synth, weights as W
    from rapiddoc_b200.engine import DetEngine, RecEngine
    from rapiddoc_b200.parallel import broadcast_blob, shard_range
    prec = PREC_FP16 if args.precision == "fp16" else PREC_FP32
    esz = 2 if prec == PREC_FP16 else 4
    # weights: rank 0 packs, NCCL broadcast over NVLink at init (the only collective on the path)
    blob = broadcast_blob((W.det_blob() if wl_key == "det" else W.rec_blob()) if rank == 0 else None, local)
    B, H, Wd = wl["batch"], wl["h"], wl["w"]
    # every rank generates ITS shard of the global synthetic batch (weak scaling: B per GPU)
    lo, hi = shard_range(world * B, world, rank)
    if wl_key == "det":
        uniq = min(B, 8)
        base = synth.det_pages(uniq, H, Wd, seed=1 + rank)
        host = torch.empty((B, H, Wd, 3), dtype=torch.uint8).pin_memory()
        for i in range(B):  # distinct pages: roll the unique ones by a per-page offset
            host[i] = torch.from_numpy(np.roll(base[i % uniq], shift=(7 * (i // uniq), 13 * (i // uniq)), axis=(0, 1)))
        eng = DetEngine(device=local, precision=prec, blob=blob)
        if args.chunk_pixels:
            eng.set_chunk_pixels(args.chunk_pixels)
        dev_in = host.cuda()
        d_prob = torch.empty((B, H, Wd), dtype=torch.float32, device="cuda")
        d_bm = torch.empty((B, H, Wd), dtype=torch.uint8, device="cuda")
        h_prob = torch.empty((B, H, Wd), dtype=torch.float32).pin_memory()
        h_bm = torch.empty((B, H, Wd), dtype=torch.uint8).pin_memory()
        host_np, h_prob_np, h_bm_np = host.numpy(), h_prob.numpy(), h_bm.numpy()

        def step_dev():
            eng.infer_u8(dev_in, prob=d_prob, bitmap=d_bm, stream=torch.cuda.current_stream())

        def step_host():
            eng.infer_u8(host_np, prob=h_prob_np, bitmap=h_bm_np)
        h2d, d2h = host.numel(), h_prob.numel() * 4 + h_bm.numel()
    else:
        base = synth.rec_crops(B, H, Wd, seed=2 + rank)
        host = torch.from_numpy(base).pin_memory()
        vw = np.full(B, Wd, np.int32)
        eng = RecEngine(device=local, precision=prec, blob=blob)
        dev_in = host.cuda()
        d_vw = torch.from_numpy(vw).cuda()
        T = eng.tokens(Wd)
        d_out = eng._outs(B, T, dev_in, False)
        h_out = {k: (v.cpu().pin_memory().numpy() if v is not None else None) for k, v in d_out.items()}
        host_np = host.numpy()

        def step_dev():
            eng.infer_u8(dev_in, d_vw, stream=torch.cuda.current_stream(), outs=d_out)

        def step_host():
            eng.infer_u8(host_np, vw, outs=h_out)
        h2d, d2h = host.numel() + B * 4, B * T * 12 + B * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident, CUDA events on the launching (torch current) stream
    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        step_dev()
        launches += eng.last_launches
    e1.record()
    barrier()
    clocks = sampler.stop(t_begin, time.time())
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * args.steps / (ms / 1e3)
    # ---- e2e: host pinned buffers through the C-ABI, copies inside the timed region
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = world * B * args.steps / dt
    # ---- roofline of the dominant kernel: profiled pass (events around every launch)
    roofline = None
    if rank == 0:
        os.environ["RDB_LANES"] = "1"     # one compute lane: kernels run back to back, so event pairs time ONE kernel each
        _lib.load().rdb_switches_reload()
        _lib.profile(True)
        _lib.profile_reset()
        for _ in range(2):
            step_dev()
        torch.cuda.synchronize()
        prof = _lib.profile_dump()
        _lib.profile(False)
        os.environ.pop("RDB_LANES", None)
        _lib.load().rdb_switches_reload()
        total = sum(v[0] for v in prof.values())
        if args.profile_out:
            rows = [{"kernel": k, "total_ms": v[0], "launches": v[1], "avg_us": v[0] / v[1] * 1e3, "share": v[0] / total,
                     "algorithmic_bytes": algorithmic_bytes(k, esz, wl, max(1, min(B, (args.chunk_pixels or 32 * 1024 * 1024) // (H * Wd))) if wl_key == "det" else B)}
                    for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])]
            for r in rows:
                if r["algorithmic_bytes"]:
                    r["gbps"] = r["algorithmic_bytes"] / (r["avg_us"] * 1e-6) / 1e9
            json.dump({"workload": wl["name"], "precision": args.precision, "steps_profiled": 2, "total_ms": total, "kernels": rows},
                      open(args.profile_out, "w"), indent=1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        n_chunk = max(1, min(B, (args.chunk_pixels or 32 * 1024 * 1024) // (H * Wd))) if wl_key == "det" else min(B, 256)
        # dominant kernel = largest share of device time among kernels that move a modelled number of bytes
        # (latency-bound helpers such as the SE FC stack have no byte model; they stay visible in top5)
        modelled = {k: v for k, v in prof.items() if algorithmic_bytes(k, esz, wl, n_chunk)}
        name, (kms, kn) = max((modelled or prof).items(), key=lambda kv: kv[1][0])
        ab = algorithmic_bytes(name, esz, wl, n_chunk)
        avg_s = kms / kn / 1e3
        ach = (ab / avg_s / 1e9) if ab else None
        traffic = None
        try:   # dram__bytes_read+write per launch of this kernel from the committed ncu --set full capture, if one exists
            for row in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))):
                if row["kernel"] == name:
                    traffic = row["dram_bytes"]
        except Exception:
            pass
        roofline = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": (ach / hbm) if ach else None,
                    "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                    "algorithmic_bytes_per_launch": ab, "avg_launch_us": avg_s * 1e6, "share_of_step": kms / total,
                    "top5": [{"kernel": k, "share": v[0] / total, "avg_us": v[0] / v[1] * 1e3} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:5]]}
        fl = algorithmic_flops(name)
        if fl:
            roofline["achieved_tflops"] = fl / avg_s / 1e12
    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = 8 if wl_key == "det" else 128
        data = host_np[:sample]
        threads = pick_cpu_threads(wl_key, data)
        v, dt_cpu = time_cpu(wl_key, data, sample, reps=2 if wl_key == "det" else 3)
        cpu = {"value": v, "unit": wl["unit"], "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
               "sample": f"{sample} {wl['unit'].split('/')[0]} of this step's batch, torch CPU fp32 oracle incl. normalise + DB bitmap / CTC decode, {dt_cpu:.1f}s per pass"}
    if rank == 0:
        line = {"metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16" if prec == PREC_FP16 else "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "batch_per_gpu": B, "h": H, "w": Wd, "precision": args.precision, "input": "uint8 BGR HWC",
                           "l2": f"inputs {host.numel() / 1e6:.0f} MB + activations per step exceed the 126 MB L2 (no explicit flush)",
                           "parallelism": f"page-parallel replicas x{world}, NCCL weight broadcast at init only"},
                "clocks": clocks, "e2e": {"value": e2e, "unit": wl["unit"], "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        emit_line(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
