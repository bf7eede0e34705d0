"""GPU suite: resident mode (DRA_CFG_RESIDENT) — the single-launch kernel stays up and takes batches by doorbell.
Same bytes as the oracle; state carries over between batches; every other entry point stops the resident kernel first;
an idle kernel leaves by itself and is restarted on demand."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_resident_batches_match_oracle(pkg, oracle):
    R, A = pkg.records, pkg.api
    w = pkg.synth.cfg2(6000, 60)
    ref, ref_inv = oracle.allocate(w.gpus, w.node_off, w.table, w.claims)
    with A.Context(device=0, flags=A.CFG_RESIDENT) as c:
        c.set_table(w.table); c.set_inventory(w.gpus, w.node_off)
        pc = A.PinnedBuffer(w.n_claim, R.CLAIM_DTYPE); pc.array[:] = w.claims
        po = A.PinnedBuffer(w.n_out, R.OUT_DTYPE)
        l0 = c.launch_count()
        for it in range(50):                                           # one launch, fifty batches
            po.array[:] = 0
            c.allocate(pc.array, None, w.n_out, flags=A.F_FRESH_INVENTORY, out=po.array)
            assert po.array.tobytes() == ref.tobytes(), it
        assert c.launch_count() - l0 == 1 and c.serve_batches() == 50
        assert c.get_inventory().tobytes() == ref_inv.tobytes()        # (stops the resident kernel, reads the state it left)
        # state accumulates without FRESH: two batches in a row = the oracle run twice
        c.reset_inventory()
        a = c.allocate(pc.array, None, w.n_out, out=po.array).copy()
        b = c.allocate(pc.array, None, w.n_out, out=po.array).copy()
        r1, inv1 = oracle.allocate(w.gpus, w.node_off, w.table, w.claims)
        r2, inv2 = oracle.allocate(inv1, w.node_off, w.table, w.claims)
        assert a.tobytes() == r1.tobytes() and b.tobytes() == r2.tobytes()
        # another entry point in between: UnsuitableNodes sees the state the resident kernel left, then the kernel comes back
        pod_off = np.arange(0, 201, dtype=np.uint32)
        bits = c.unsuitable(w.claims[:200], pod_off)
        cn = np.tile(np.arange(w.n_node, dtype=np.uint32), 200); co = (np.arange(201, dtype=np.uint32) * w.n_node).astype(np.uint32)
        assert bits.tobytes() == oracle.unsuitable(inv2, w.node_off, w.table, w.claims[:200], pod_off, cn, co).tobytes()
        r3, _ = oracle.allocate(inv2, w.node_off, w.table, w.claims)
        assert c.allocate(pc.array, None, w.n_out, out=po.array).tobytes() == r3.tobytes()
        # different shapes through the same resident kernel: smaller batch, out_off with counts > 1, malformed claims
        m = pkg.synth.mixed(3000, 60, 9)
        c.set_inventory(m.gpus, m.node_off)
        pm = A.PinnedBuffer(m.n_claim, R.CLAIM_DTYPE); pm.array[:] = m.claims
        pf = A.PinnedBuffer(m.n_claim, np.uint32); pf.array[:] = m.out_off
        pq = A.PinnedBuffer(m.n_out, R.OUT_DTYPE)
        rm, invm = oracle.allocate(m.gpus, m.node_off, m.table, m.claims, m.out_off, m.n_out)
        for _ in range(3):
            got = c.allocate(pm.array, pf.array, m.n_out, flags=A.F_FRESH_INVENTORY, out=pq.array)
            assert got.tobytes() == rm.tobytes()
        assert c.get_inventory().tobytes() == invm.tobytes()
        # the kernel leaves after DRA_SERVE_IDLE_MS (20 ms) without a doorbell and is started again by the next call
        c.allocate(pm.array, pf.array, m.n_out, flags=A.F_FRESH_INVENTORY, out=pq.array)
        l1 = c.launch_count()
        time.sleep(0.2)
        got = c.allocate(pm.array, pf.array, m.n_out, flags=A.F_FRESH_INVENTORY, out=pq.array)
        assert got.tobytes() == rm.tobytes() and c.launch_count() == l1 + 1
        # a batch that does not fit the staged form stops the kernel and takes the usual path
        big = pkg.synth.cfg2(30000, 60)
        rb, _ = oracle.allocate(big.gpus, big.node_off, big.table, big.claims)
        c.set_inventory(big.gpus, big.node_off)
        assert c.allocate(big.claims, flags=A.F_FRESH_INVENTORY).tobytes() == rb.tobytes()
        for p in (pc, po, pm, pf, pq):
            p.free()


def test_resident_kernel_is_stopped_by_destroy(pkg, oracle):
    R, A = pkg.records, pkg.api
    w = pkg.synth.cfg2(2000, 20)
    ref, _ = oracle.allocate(w.gpus, w.node_off, w.table, w.claims)
    c = A.Context(device=0, flags=A.CFG_RESIDENT)
    c.set_table(w.table); c.set_inventory(w.gpus, w.node_off)
    pc = A.PinnedBuffer(w.n_claim, R.CLAIM_DTYPE); pc.array[:] = w.claims
    po = A.PinnedBuffer(w.n_out, R.OUT_DTYPE)
    assert c.allocate(pc.array, None, w.n_out, flags=A.F_FRESH_INVENTORY, out=po.array).tobytes() == ref.tobytes()
    t0 = time.perf_counter()
    c.serve_stop()                                                     # EXIT command: the kernel leaves at once, not after the idle time-out
    assert time.perf_counter() - t0 < 0.01
    c.close()
    assert time.perf_counter() - t0 < 0.1                                # (the idle time-out alone would be 20 ms + teardown)
    with A.Context(device=0) as d:                                     # the device is free again
        d.set_table(w.table); d.set_inventory(w.gpus, w.node_off)
        assert d.allocate(w.claims).tobytes() == ref.tobytes()
    pc.free(); po.free()


def test_calibration_keeps_state_and_results(pkg, oracle):
    """dra_calibrate measures the single-launch / sort-path crossover on this box; it must leave the inventory alone and
    the answers unchanged on both sides of whatever crossover it finds."""
    w = pkg.synth.cfg2(9000, 60)
    first, inv1 = oracle.allocate(w.gpus, w.node_off, w.table, w.claims[:3000])
    with pkg.api.Context(device=0) as c:
        c.set_table(w.table); c.set_inventory(w.gpus, w.node_off)
        assert c.allocate(w.claims[:3000].copy()).tobytes() == first.tobytes()
        x = c.calibrate()
        print("single-launch kernel beats the sort path up to", x, "claims on this box (60 nodes)")
        assert 0 <= x <= 24000
        assert c.get_inventory().tobytes() == inv1.tobytes()
        for n in (500, 5000, 9000):
            ref, _ = oracle.allocate(inv1, w.node_off, w.table, w.claims[:n])
            c.set_inventory(inv1, w.node_off)
            assert c.allocate(w.claims[:n].copy()).tobytes() == ref.tobytes()
