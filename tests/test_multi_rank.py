"""CPU suite, part 3: the N>1 host logic (node sharding, local renumbering, merge after the all-gather)
with world_size 2 and 3 over gloo.  The per-rank allocation is done by the CPU oracle here — it stands in
for the GPU only inside this test; on the GPU box bench.py --gpus N runs the real thing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    pkg = importlib.import_module("k8s-dra-driver_b200")
    from oracle import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = pkg.synth.mixed(3000, 31, 5)                    # every rank builds the same global batch
        ranges = pkg.shard.plan(w.claims["node"], w.n_node, world)
        lb = pkg.shard.local_batch(w.gpus, w.node_off, w.claims, rank, ranges)
        out, inv = O.allocate(lb.gpus, lb.node_off, w.table, lb.claims, lb.out_off, lb.n_out)
        # the path's one collective: all-gather of (padded) OutRec slices
        n_per = torch.tensor([lb.n_out], dtype=torch.int64)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, n_per)
        pad = int(max(int(s.item()) for s in sizes))
        mine = np.zeros(pad, dtype=pkg.records.OUT_DTYPE)
        mine[: lb.n_out] = out
        t = torch.from_numpy(mine.view(np.uint8).copy())
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        parts = []
        for r in range(world):
            lr = pkg.shard.local_batch(w.gpus, w.node_off, w.claims, r, ranges)     # index arithmetic only
            parts.append((lr.sel, lr.out_off, gathered[r].numpy().view(pkg.records.OUT_DTYPE)[: lr.n_out], lr.gpu_base))
        merged = pkg.shard.merge(w.n_out, w.out_off, parts)
        ref, ref_inv = O.allocate(w.gpus, w.node_off, w.table, w.claims, w.out_off, w.n_out)
        ok = merged.tobytes() == ref.tobytes()
        g0, g1 = int(w.node_off[lb.n0]), int(w.node_off[lb.n1])
        inv_g = inv.copy(); inv_g["node"] += np.uint32(lb.n0)
        ok = ok and inv_g.tobytes() == ref_inv[g0:g1].tobytes()
        q.put((rank, ok, ranges))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_allocation_matches_single_rank(world, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    ranges = res[0][2]
    assert ranges[0][0] == 0 and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


def test_plan_is_balanced_and_contiguous(pkg):
    w = pkg.synth.cfg3(20_000, 200)
    for world in (1, 2, 4, 8):
        r = pkg.shard.plan(w.claims["node"], w.n_node, world)
        assert r[0][0] == 0 and r[-1][1] == w.n_node
        loads = [int(((w.claims["node"] >= a) & (w.claims["node"] < b)).sum()) for a, b in r]
        assert sum(loads) == w.n_claim
        assert max(loads) - min(loads) <= 2 * (w.n_claim // w.n_node + 50)


# ---- the GLOBAL batch form (round 2): every rank holds the whole inventory and the whole claim array -------------------
def _global_worker(rank, world, port, q):
    """What dra_allocate_batch_global_device does, restated with the CPU oracle standing in for the GPU: keep the claims of
    this rank's node range IN INPUT ORDER (stable compaction), allocate them on the global inventory, and all-gather the
    records as (global slot, OutRec) packets; every rank assembles the table from its own records + the packets."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    pkg = importlib.import_module("k8s-dra-driver_b200")
    from oracle import oracle as O
    R = pkg.records
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = pkg.synth.mixed(3000, 31, 6)
        ranges = pkg.shard.plan(w.claims["node"], w.n_node, world)
        lo, hi = ranges[rank]
        node = w.claims["node"]
        keep = ((node >= lo) & (node < hi)) | ((node >= w.n_node) & (rank == 0))      # stray claims: rank 0
        idx = np.nonzero(keep)[0]
        mine = w.claims[idx].copy()
        slots = R.claim_slots(w.claims, w.n_node)
        goff = w.out_off                                                            # GLOBAL first slot of every claim
        loff = np.zeros(len(idx), np.uint32); loff[1:] = np.cumsum(slots[idx][:-1])
        out, inv = O.allocate(w.gpus, w.node_off, w.table, mine, loff, int(slots[idx].sum()))
        # packets: (global slot, record) for every slot this rank answers
        gs = np.concatenate([np.arange(goff[i], goff[i] + slots[i]) for i in idx]).astype(np.int64) if len(idx) else np.zeros(0, np.int64)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([len(gs)], dtype=torch.int64))
        pad = int(max(int(s_.item()) for s_ in sizes))
        buf = np.zeros((pad, 2), dtype=np.int64)
        buf[: len(gs), 0] = gs; buf[: len(gs), 1] = out.view(np.uint64).astype(np.int64)
        gathered = [torch.zeros(pad, 2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(buf))
        table = np.zeros(w.n_out, dtype=np.uint64)
        for r in range(world):
            n = int(sizes[r].item())
            g = gathered[r].numpy()
            table[g[:n, 0]] = g[:n, 1].astype(np.uint64)
        ref, ref_inv = O.allocate(w.gpus, w.node_off, w.table, w.claims, w.out_off, w.n_out)
        ok = table.tobytes() == ref.tobytes()
        g0, g1 = int(w.node_off[lo]), int(w.node_off[hi])
        ok = ok and inv[g0:g1].tobytes() == ref_inv[g0:g1].tobytes()               # a rank's inventory is authoritative for its range
        q.put((rank, ok, sum(int(s_.item()) for s_ in sizes) == w.n_out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_global_batch_sharded_by_node_range(world, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_global_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and covered for _, ok, covered in res)       # same bytes as one rank, and every slot answered exactly once
