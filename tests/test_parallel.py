"""CPU tests of the N>1 host logic with the gloo backend, world_size 2."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from rapiddoc_b200 import parallel as PL


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 1200):
        for world in (1, 2, 3, 8):
            spans = [PL.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_deal_sorted_by_width_is_a_permutation():
    r = np.random.default_rng(0).uniform(1, 30, 101)
    parts = PL.deal_sorted_by_width(r, 4)
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(101))
    means = [r[p].mean() for p in parts]
    assert max(means) - min(means) < 1.5


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rapiddoc_b200 import weights as W
    blob = W.pack({"a.w": np.arange(24, dtype=np.float32).reshape(2, 3, 4), "b": np.ones(5, np.float32)}) if rank == 0 else None
    got = PL.broadcast_blob(blob)
    lo, hi = PL.shard_range(11, world, rank)
    res = PL.gather_objects({"rank": rank, "units": list(range(lo, hi))})
    q.put((rank, len(got), got[:4], sum(len(r["units"]) for r in res), sorted(sum((r["units"] for r in res), []))))
    dist.destroy_process_group()


def test_gloo_world2_broadcast_and_gather():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(30) for p in ps]
    assert out[0][1] == out[1][1] > 0 and out[0][2] == out[1][2] == b"RDW1"
    assert out[0][3] == out[1][3] == 11 and out[0][4] == list(range(11))


def test_pin_rank_to_cores_partitions_the_affinity_mask():
    import os
    from rapiddoc_b200.parallel import pin_rank_to_cores
    if not hasattr(os, "sched_getaffinity"):
        return
    before = sorted(os.sched_getaffinity(0))
    try:
        shares = []
        for r in range(2):
            os.sched_setaffinity(0, before)
            shares.append(pin_rank_to_cores(r, 2, ideal=None))
        if len(before) >= 2:
            assert not set(shares[0]) & set(shares[1]) and len(shares[0]) == len(shares[1]) == len(before) // 2
            from rapiddoc_b200 import dbpost
            assert dbpost.host_threads() == len(shares[1])
    finally:
        os.sched_setaffinity(0, before)


def test_rank_core_plan_follows_gpu_numa_locality():
    from rapiddoc_b200.parallel import plan_rank_cores
    cores = list(range(32))
    assert plan_rank_cores(cores, 4) == [list(range(0, 8)), list(range(8, 16)), list(range(16, 24)), list(range(24, 32))]
    # 8 GPUs, two NUMA nodes with interleaved core numbering (even cores node 0, odd cores node 1); GPUs 0-3 on node 0
    node0, node1 = set(range(0, 32, 2)), set(range(1, 32, 2))
    plan = plan_rank_cores(cores, 8, [node0] * 4 + [node1] * 4)
    assert all(len(p) == 4 for p in plan) and all(set(p) <= node0 for p in plan[:4]) and all(set(p) <= node1 for p in plan[4:])
    assert len({c for p in plan for c in p}) == 32                                   # disjoint, all cores used
    # the job may only use 8 cores, all on node 0: ranks on node 1 fall back to their flat share
    plan = plan_rank_cores(list(range(0, 16, 2)), 4, [node0, node0, node1, node1])
    assert plan[0] == [0, 2, 4, 6][:len(plan[0])] and plan[2] == [8, 10] and plan[3] == [12, 14]
    # one GPU with every core near it: capped at twice the flat share is irrelevant for a single rank
    assert plan_rank_cores(cores, 1, [set(cores)]) == [cores]
    assert plan_rank_cores([0, 1], 4) == [[0, 1]] * 4                                # fewer cores than ranks: no pinning split


def test_smt_order_keeps_siblings_adjacent():
    import os
    from rapiddoc_b200.parallel import plan_rank_cores, smt_order
    if not hasattr(os, "sched_getaffinity"):
        return
    cores = sorted(os.sched_getaffinity(0))
    order = smt_order(cores)
    assert sorted(order) == cores
    # a 16-core / 32-thread box numbered the Linux way: cpu c and c+16 share a core -> ranks get whole physical cores
    fake = [c for pair in zip(range(16), range(16, 32)) for c in pair]
    plan = plan_rank_cores(fake, 4)
    assert plan[0] == [0, 16, 1, 17, 2, 18, 3, 19] and plan[3] == [12, 28, 13, 29, 14, 30, 15, 31]
