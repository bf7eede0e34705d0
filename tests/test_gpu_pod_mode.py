"""GPU suite: pod mode and the exhaustive placement search (spec §12) through the C ABI, bit-exact with the oracle."""
import os

import numpy as np
import pytest

from test_gpu_parity import _assert_same

pytestmark = pytest.mark.gpu


def _dense(n_pod, n_node):
    return (np.tile(np.arange(n_node, dtype=np.uint32), n_pod),
            (np.arange(n_pod + 1, dtype=np.uint32) * n_node).astype(np.uint32))


def test_order_dependence_example(pkg, ctx, oracle):
    """VERDICT r01's counter-example: slice 4 busy, pod {1g, 3g}."""
    R = pkg.records
    g, off = R.make_inventory([1], mig=True, busy=1 << 4)
    t = R.default_table()
    ctx.set_table(t); ctx.set_inventory(g, off)
    c = np.zeros(2, dtype=R.CLAIM_DTYPE)
    c["kind"], c["count"], c["profile"] = R.KIND_MIG, 1, [R.GI_1_SLICE, R.GI_3_SLICE]
    pod_off = np.array([0, 2], np.uint32)
    assert ctx.unsuitable(c, pod_off)[0] & 1 == 0                          # first-fit: false negative
    assert ctx.unsuitable(c, pod_off, flags=pkg.api.F_EXHAUSTIVE)[0] & 1 == 1
    out = ctx.allocate_pods(c, pod_off)
    assert list(out["status"]) == [R.ST_POD, R.ST_POD] and ctx.get_inventory().tobytes() == g.tobytes()
    out = ctx.allocate_pods(c, pod_off, flags=pkg.api.F_EXHAUSTIVE)
    assert list(out["status"]) == [0, 0] and list(out["start"]) == [5, 0]
    ref, inv = oracle.allocate_pods(g, off, t, c, pod_off, flags=oracle.F_EXHAUSTIVE)
    _assert_same(out, ctx.get_inventory(), ref, inv, "judge example")
    # 7 x 1g + 7g on one empty GPU: the budget runs out on the device exactly as in the oracle
    g, off = R.make_inventory([1], mig=True)
    ctx.set_inventory(g, off)
    c = np.zeros(8, dtype=R.CLAIM_DTYPE)
    c["kind"], c["count"], c["profile"] = R.KIND_MIG, 1, [R.GI_1_SLICE] * 7 + [R.GI_7_SLICE]
    out = ctx.allocate_pods(c, np.array([0, 8], np.uint32), flags=pkg.api.F_EXHAUSTIVE)
    assert set(out["status"]) == {R.ST_SEARCH_LIMIT} and ctx.get_inventory().tobytes() == g.tobytes()


@pytest.mark.parametrize("seed,n_pod,n_node,width", [(0, 3000, 16, 8), (1, 20000, 64, 8), (2, 5000, 9, 16), (3, 4000, 5, 32),
                                                     (4, 60000, 300, 8)])
def test_pods_match_oracle(pkg, ctx, oracle, seed, n_pod, n_node, width):
    w, pod_off = pkg.synth.pods(n_pod, n_node, seed, gpus_per_node=width)
    for flags, oflags in ((0, 0), (pkg.api.F_EXHAUSTIVE, oracle.F_EXHAUSTIVE)):
        ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
        out = ctx.allocate_pods(w.claims, pod_off, flags=flags)
        ref, inv = oracle.allocate_pods(w.gpus, w.node_off, w.table, w.claims, pod_off, flags=oflags)
        _assert_same(out, ctx.get_inventory(), ref, inv, f"pods seed {seed} flags {flags}")
        assert (out["status"] == 0).any() and (out["status"] == pkg.records.ST_POD).any()
    # UnsuitableNodes, dense, on the ORIGINAL inventory: first-fit vs exhaustive
    ctx.set_inventory(w.gpus, w.node_off)
    sub = min(n_pod, 2000)
    po = pod_off[: sub + 1]
    c = w.claims[: po[-1]]
    cn, co = _dense(sub, n_node)
    ex = ctx.unsuitable(c, po, flags=pkg.api.F_EXHAUSTIVE)
    ref = oracle.unsuitable(w.gpus, w.node_off, w.table, c, po, cn, co, flags=oracle.F_EXHAUSTIVE)
    assert ex[: len(ref)].tobytes() == ref.tobytes()
    ff = np.unpackbits(ctx.unsuitable(c, po), bitorder="little")[: sub * n_node]
    exb = np.unpackbits(ex, bitorder="little")[: sub * n_node]
    ungrouped = np.repeat(np.add.reduceat(c["group"], po[:-1].astype(np.int64)) == 0, n_node)
    assert not (ff & ~exb & ungrouped).any()                               # exhaustive never loses a first-fit fit (same semantics w/o groups)
    assert (exb & ~ff).any()                                               # and finds fits first-fit misses


def test_pods_with_selectors_and_mixed_kinds(pkg, ctx, oracle):
    """Pods cut out of the 'mixed' workload: GPU / SHARED / MIG claims, counts > 1, malformed claims, heterogeneous
    ragged nodes, per-claim selectors."""
    R = pkg.records
    for seed in range(4):
        w = pkg.synth.mixed(3000, 23, 40 + seed, invalid=seed % 2 == 0)
        attrs, sels = (None, None)
        if seed >= 2:
            attrs, sels = pkg.synth.with_selectors(w, seed)
        rng = np.random.default_rng(seed)
        cuts = np.unique(np.concatenate([[0, w.n_claim], rng.integers(0, w.n_claim, 900)])).astype(np.uint32)
        c = w.claims.copy()
        for p in range(len(cuts) - 1):
            if rng.integers(0, 12):
                c["node"][cuts[p]: cuts[p + 1]] = c["node"][cuts[p]]
        out_off, n_out = R.out_offsets(c, w.n_node)
        ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
        ctx.set_gpu_attrs(attrs); ctx.set_selectors(None if sels is None else sels.reshape(-1, R.SEL_MAX_INS))
        oracle.set_selectors(attrs, sels)
        try:
            for flags, oflags in ((0, 0), (pkg.api.F_EXHAUSTIVE, oracle.F_EXHAUSTIVE)):
                ctx.reset_inventory()
                out = ctx.allocate_pods(c, cuts, out_off, n_out, flags)
                ref, inv = oracle.allocate_pods(w.gpus, w.node_off, w.table, c, cuts, out_off, n_out, oflags)
                _assert_same(out, ctx.get_inventory(), ref, inv, f"mixed pods seed {seed} flags {flags}")
                ctx.deallocate(c, out, out_off)
                assert ctx.get_inventory().tobytes() == w.gpus.tobytes()
        finally:
            oracle.set_selectors()
            ctx.set_selectors(None); ctx.set_gpu_attrs(None)


def test_pods_on_arbitrary_problems(pkg, ctx, oracle):
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st
    from test_pod_mode import pod_problems
    n_ex = int(os.environ.get("DRA_PROP_EXAMPLES", "0"))

    @settings(max_examples=n_ex or 120, deadline=None, suppress_health_check=list(HealthCheck), derandomize=n_ex == 0, database=None)
    @given(pod_problems(), st.booleans())
    def run(prob, exhaustive):
        g, off, t, c, pod_off, out_off, n_out = prob
        flags, oflags = (pkg.api.F_EXHAUSTIVE, oracle.F_EXHAUSTIVE) if exhaustive else (0, 0)
        ctx.set_table(t); ctx.set_inventory(g, off)
        n_pod, n_node = len(pod_off) - 1, len(off) - 1
        if len(c) and exhaustive:
            bits = ctx.unsuitable(c, pod_off, flags=flags)
            cn, co = _dense(n_pod, n_node)
            ref = oracle.unsuitable(g, off, t, c, pod_off, cn, co, flags=oflags)
            assert bits[: len(ref)].tobytes() == ref.tobytes(), "exhaustive UnsuitableNodes differs"
        out = ctx.allocate_pods(c, pod_off, out_off, n_out, flags)
        ref, inv = oracle.allocate_pods(g, off, t, c, pod_off, out_off, n_out, oflags)
        _assert_same(out, ctx.get_inventory(), ref, inv, "arbitrary pods")

    run()
