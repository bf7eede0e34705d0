"""GPU suite: the multi-rank paths on ONE device — world contexts of this process, each with its own stream, wired
together with dra_comm_init_local / dra_peer_import_local (the form a single driver process uses; torchrun ranks use
IPC handles instead: tests/multi_gpu_check.py, wrapped by test_gpu_multi_rank.py).  Covers the sharded GLOBAL batch
(device-side range filter, results at global slots, packet all-gather) and the per-rank-slice gather, on the fused and
the sort path, bit-exact with the oracle on the global batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy()).to("cuda:0")


def _make_world(pkg, world, flags=0):
    ctxs = [pkg.api.Context(device=0, flags=flags) for _ in range(world)]
    for r, c in enumerate(ctxs):
        c.comm_init_local(r, world)
    return ctxs


@pytest.mark.parametrize("world,cfg_flags,seed", [(2, 0, 0), (4, 0, 1), (3, "nofused", 2), (2, 0, 3), (2, "inkernel", 4), (4, "inkernel", 1)])
def test_sharded_global_batch(pkg, oracle, world, cfg_flags, seed, monkeypatch):
    R = pkg.records
    if cfg_flags == "inkernel":                                          # experiment: the compaction INSIDE k_fused (one launch per call)
        monkeypatch.setenv("DRA_SHARD_IN_KERNEL", "1")
    flags = pkg.api.CFG_NO_FUSED if cfg_flags == "nofused" else 0
    w = pkg.synth.mixed(6000, 30, 60 + seed, invalid=seed != 3) if seed != 1 else pkg.synth.cfg2(8000, 40)
    out_off = w.out_off
    ranges = pkg.shard.plan(w.claims["node"], w.n_node, world)          # one-time setup, like the inventory
    ctxs = _make_world(pkg, world, flags)
    try:
        for r, c in enumerate(ctxs):
            c.set_table(w.table); c.set_inventory(w.gpus, w.node_off)
            if seed == 3:
                c.set_shard(ranges[r][0], ranges[r][1], take_stray=(r == 0))      # without the map: counts travel in a header
            else:
                c.set_shard_map([x[0] for x in ranges] + [ranges[-1][1]], stray_rank=0)
            c.shard_export(w.n_out, 0, want_handle=False)
        for c in ctxs:
            c.peer_import_local(ctxs)
        d_claims = _dev(w.claims)
        d_off = None if out_off is None else _dev(out_off)
        ref, ref_inv = oracle.allocate(w.gpus, w.node_off, w.table, w.claims, out_off, w.n_out)
        for rep in range(4):                                             # repeated calls: parity double-buffering, plan from the hint
            l0 = ctxs[0].launch_count()
            for c in ctxs:                                               # enqueue on every rank, THEN wait
                c.allocate_global_device(d_claims.data_ptr(), w.n_claim, None if d_off is None else d_off.data_ptr(), w.n_out,
                                         pkg.api.F_FRESH_INVENTORY)
            if cfg_flags == "inkernel" and rep:
                assert ctxs[0].launch_count() - l0 == 1, "the in-kernel compaction was not taken"
            for r, c in enumerate(ctxs):
                got = c.gather_read(np.zeros(w.n_out, dtype=R.OUT_DTYPE))
                assert got.tobytes() == ref.tobytes(), f"rank {r} rep {rep}: table differs from the oracle"
                g0, g1 = int(w.node_off[ranges[r][0]]), int(w.node_off[ranges[r][1]])
                assert c.get_inventory()[g0:g1].tobytes() == ref_inv[g0:g1].tobytes(), f"rank {r}: shard inventory differs"
    finally:
        for c in ctxs:
            c.close()


def test_sharded_global_batch_world1_and_empty_shards(pkg, oracle):
    """world 1 needs no export; a rank whose node range gets no claim; a rank with an empty node range."""
    R = pkg.records
    w = pkg.synth.cfg2(3000, 12)
    w.claims["node"] = w.claims["node"] % 5                              # nodes 5..11 get nothing
    ref, _ = oracle.allocate(w.gpus, w.node_off, w.table, w.claims)
    with pkg.api.Context(device=0) as c:
        c.set_table(w.table); c.set_inventory(w.gpus, w.node_off); c.set_shard(0, w.n_node, True)
        d = _dev(w.claims)
        c.allocate_global_device(d.data_ptr(), w.n_claim, None, w.n_out, pkg.api.F_FRESH_INVENTORY)
        assert c.gather_read(np.zeros(w.n_out, dtype=R.OUT_DTYPE)).tobytes() == ref.tobytes()
    ctxs = _make_world(pkg, 3)
    try:
        rng = [(0, 5), (5, 12), (12, 12)]
        for r, c in enumerate(ctxs):
            c.set_table(w.table); c.set_inventory(w.gpus, w.node_off); c.set_shard_map([0, 5, 12, 12], stray_rank=2)
            c.shard_export(w.n_out, 0, want_handle=False)
        for c in ctxs:
            c.peer_import_local(ctxs)
        d = _dev(w.claims)
        for _ in range(3):
            for c in ctxs:
                c.allocate_global_device(d.data_ptr(), w.n_claim, None, w.n_out, pkg.api.F_FRESH_INVENTORY)
            for c in ctxs:
                assert c.gather_read(np.zeros(w.n_out, dtype=R.OUT_DTYPE)).tobytes() == ref.tobytes()
        # the distribution shifts under a plan made for the old one: every claim now lands in rank 1's range.  The call
        # fails WITHOUT touching anything ("call again"), the peers are told at once (no time-out), the retry succeeds.
        w2 = pkg.synth.cfg2(3000, 12); w2.claims["node"] = 5 + w2.claims["node"] % 7
        ref2, _ = oracle.allocate(w2.gpus, w2.node_off, w2.table, w2.claims)
        d2 = _dev(w2.claims)
        for c in ctxs:
            c.allocate_global_device(d2.data_ptr(), w2.n_claim, None, w2.n_out, pkg.api.F_FRESH_INVENTORY)
        errs = []
        for c in ctxs:
            try:
                c.gather_read(np.zeros(w2.n_out, dtype=R.OUT_DTYPE))
            except pkg.api.DraError as e:
                errs.append(e.code)
        assert pkg.api.E_STATE in errs, errs
        for c in ctxs:
            c.allocate_global_device(d2.data_ptr(), w2.n_claim, None, w2.n_out, pkg.api.F_FRESH_INVENTORY)
        for c in ctxs:
            assert c.gather_read(np.zeros(w2.n_out, dtype=R.OUT_DTYPE)).tobytes() == ref2.tobytes()
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("world,nofused", [(2, False), (3, True)])
def test_per_rank_slices_gather(pkg, oracle, world, nofused):
    """dra_allocate_batch_gather_device: every rank owns its nodes' inventory and claims; slices + zeroed padding."""
    R = pkg.records
    w = pkg.synth.mixed(5000, 28, 77)
    ranges = pkg.shard.plan(w.claims["node"], w.n_node, world)
    lbs = [pkg.shard.local_batch(w.gpus, w.node_off, w.claims, r, ranges) for r in range(world)]
    n_per = max(b.n_out for b in lbs) + 3
    ctxs = _make_world(pkg, world, pkg.api.CFG_NO_FUSED if nofused else 0)
    try:
        for c, lb in zip(ctxs, lbs):
            c.set_table(w.table); c.set_inventory(lb.gpus, lb.node_off)
            c._check(c._lib.dra_peer_export(c._h, n_per, (__import__("ctypes").c_uint8 * 64)()))
        for c in ctxs:
            c.peer_import_local(ctxs)
        dc = [_dev(lb.claims) for lb in lbs]
        do = [_dev(lb.out_off) for lb in lbs]
        ref, _ = oracle.allocate(w.gpus, w.node_off, w.table, w.claims, w.out_off, w.n_out)
        for rep in range(3):
            for c, lb, a, b in zip(ctxs, lbs, dc, do):
                c.allocate_gather_device(a.data_ptr(), len(lb.claims), b.data_ptr(), None, lb.n_out, n_per, pkg.api.F_FRESH_INVENTORY)
            for c in ctxs:
                table = c.gather_read(np.zeros(world * n_per, dtype=R.OUT_DTYPE))
                parts = [(b.sel, b.out_off, table[r * n_per: r * n_per + b.n_out], b.gpu_base) for r, b in enumerate(lbs)]
                merged = pkg.shard.merge(w.n_out, w.out_off, parts)
                assert merged.tobytes() == ref.tobytes()
                for r, b in enumerate(lbs):                              # padding of every slice is zero (header promise)
                    assert not table[r * n_per + b.n_out: (r + 1) * n_per].view(np.uint64).any()
    finally:
        for c in ctxs:
            c.close()


def test_device_rendezvous_of_world_contexts(pkg):
    """dra_peer_rendezvous_device: every rank enqueues it, nobody waits on the host in between; repeated calls use
    increasing sequence numbers (nothing to reset); world 1 is a no-op."""
    import torch
    w = pkg.synth.cfg2(2000, 8)
    ctxs = _make_world(pkg, 3)
    try:
        for c in ctxs:
            c.set_table(w.table); c.set_inventory(w.gpus, w.node_off); c.set_shard_map([0, 3, 6, 8], stray_rank=0)
            c.shard_export(w.n_out, 0, want_handle=False)
        for c in ctxs:
            c.peer_import_local(ctxs)
        for _ in range(5):
            l0 = ctxs[1].launch_count()
            for c in ctxs:
                c.peer_rendezvous()
            assert ctxs[1].launch_count() == l0       # (no part of a batch: not counted)
        for c in ctxs:
            c.sync()
        torch.cuda.synchronize()
    finally:
        for c in ctxs:
            c.close()
    with pkg.api.Context(device=0) as c:
        c.peer_rendezvous()                      # world 1: nothing to do


# (A LARGE sharded batch — wave-sized hist tiles, row-wise scan, stand-alone gather — is checked on real ranks only:
#  tests/multi_gpu_check.py, "1M claims x 10k nodes".  Two contexts on ONE device cannot run it: a rank's gather kernel holds
#  every SM at the default shared-memory split while it waits for its peer, whose histogram CTAs need the SMs re-split.)
