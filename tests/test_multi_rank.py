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
