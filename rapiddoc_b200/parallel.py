"""Multi-GPU plumbing: one process per GPU, page-/crop-parallel replicas.

The hot path shards by independent units (pages, text-line crops):
there is NO data-path collective.  The only collective is the init-time broadcast of the
packed weight blob from rank 0 (NCCL over NVLink/NVSwitch on GPUs, gloo in CPU tests);
results are small per-unit structs that stay on the rank that produced them or are
gathered by the Python driver.
"""
import numpy as np


def shard_range(total: int, world: int, rank: int):
    """Contiguous balanced [lo, hi) of `total` units for `rank` (first `total % world` ranks get one extra)."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def round_robin(total: int, world: int, rank: int):
    """Indices rank handles when units are dealt round-robin (pages of a window)."""
    return list(range(rank, int(total), int(world)))


def deal_sorted_by_width(ratios, world: int):
    """Crops sorted by aspect ratio (rapid_doc/model/ocr/rapid_ocr.py:411-414) are dealt in
    contiguous runs round-robin so every rank sees a similar width mix.  Returns a list of
    index arrays, one per rank; the union is a permutation of range(len(ratios))."""
    order = np.argsort(np.asarray(ratios, dtype=np.float64), kind="stable")
    return [order[r::world] for r in range(world)]


def broadcast_blob(blob, device_index=0):
    """rank 0 passes the packed weight bytes, every rank returns them.  No-op when
    torch.distributed is not initialised (single GPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert blob is not None
        return blob
    rank = dist.get_rank()
    cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", device_index) if cuda else torch.device("cpu")
    n = torch.tensor([len(blob) if rank == 0 else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src=0)
    if rank == 0:
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    else:
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    return t.cpu().numpy().tobytes()


def gather_objects(obj):
    """Gather small per-rank Python result structs on every rank (host side, not on the data path)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def gpu_ideal_cores(local_world: int):
    """Per local GPU, the CPU set NVML reports as closest to it (its NUMA node / PCIe root), or None when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        import os
        ncpu = (max(os.sched_getaffinity(0)) // 64) + 1
        out = []
        for g in range(local_world):
            h = pynvml.nvmlDeviceGetHandleByIndex(g)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, max(ncpu, 1))
            out.append({64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1})
        return out
    except Exception:
        return None


def smt_order(cores):
    """The allowed logical CPUs ordered so that hyper-thread siblings are adjacent (physical core by physical core, from
    /sys/devices/system/cpu/cpuN/topology/thread_siblings_list): a contiguous split then hands WHOLE physical cores to a rank
    instead of giving two ranks the two halves of the same cores (Linux numbers the second hardware threads after all the
    first ones).  Falls back to numeric order."""
    cores = sorted(cores)
    allowed, groups, seen = set(cores), [], set()
    try:
        for c in cores:
            if c in seen:
                continue
            with open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list") as f:
                txt = f.read().strip()
            sib = set()
            for part in txt.split(","):
                lo, _, hi = part.partition("-")
                sib.update(range(int(lo), int(hi or lo) + 1))
            g = sorted(sib & allowed) or [c]
            seen.update(g)
            groups.append(g)
    except Exception:
        return cores
    return [c for g in sorted(groups, key=lambda g: g[0]) for c in g]


def plan_rank_cores(cores, local_world: int, ideal=None):
    """Core lists for local ranks 0..local_world-1.  With `ideal` (per-GPU nearest-CPU sets) every rank gets an equal share of
    the allowed cores NEAR ITS GPU — ranks whose GPUs hang off the same NUMA node split that node's cores; ranks whose ideal set
    is empty after intersecting with the allowed cores, and the no-NVML case, fall back to contiguous equal shares."""
    cores = list(cores)                                   # caller's order (smt_order keeps hyper-thread siblings adjacent)
    per = len(cores) // max(1, local_world)
    if per < 1:
        return [list(cores) for _ in range(local_world)]
    flat = [cores[r * per:(r + 1) * per] for r in range(local_world)]
    if not ideal or len(ideal) < local_world:
        return flat
    groups = {}
    for r in range(local_world):
        near = set(ideal[r])
        key = tuple(c for c in cores if c in near)
        groups.setdefault(key, []).append(r)
    plan = [None] * local_world
    for key, ranks in groups.items():
        share = len(key) // len(ranks)
        if share < 1:
            for r in ranks:
                plan[r] = flat[r]
            continue
        share = min(share, max(per, 1) * 2)           # never starve the other node: at most twice the flat share
        for i, r in enumerate(ranks):
            plan[r] = list(key[i * share:(i + 1) * share])
    return plan


def pin_rank_to_cores(local_rank: int, local_world: int, ideal="nvml"):
    """One process per GPU on one box: give every rank an equal share of the cores this job may use, taken from the CPUs NVML
    reports as nearest to the rank's GPU (pinned staging buffers are then first-touched on that NUMA node, and the copy
    engines do not cross the socket interconnect); contiguous equal shares when NVML gives nothing.  The host half of the
    pipeline (contours, Clipper, staging copies) sizes its thread pools from the resulting affinity mask instead of
    oversubscribing the whole machine N times.  Returns the core list (or None where affinity is not supported)."""
    import os
    try:
        cores = smt_order(os.sched_getaffinity(0))
    except Exception:
        return None
    sets = gpu_ideal_cores(local_world) if ideal == "nvml" else ideal
    mine = plan_rank_cores(cores, local_world, sets)[local_rank]
    if mine:
        os.sched_setaffinity(0, mine)
    return mine
