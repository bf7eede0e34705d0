"""CPU suite: property-based pinning of the oracle.  hypothesis draws arbitrary small problems — random placement
tables, ragged inventories with pre-occupied / blocked GPUs, claims of every kind including malformed ones and
co-location runs — and checks (1) the C oracle against the independent pure-Python restatement (oracle/naive.py),
(2) invariants spec/ALLOCATION.md states: placements come from the table and never overlap (deviceinfo.go:199-204:
occupancy is by memory slice), the inventory after the batch is exactly the inventory before plus the successful
results, and Deallocate (spec §9) of everything allocated restores it."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st


def _dicts(a):
    return [{k: int(x[k]) for k in a.dtype.names} for x in a]


@st.composite
def problems(draw):
    import importlib
    R = importlib.import_module("k8s-dra-driver_b200").records
    n_node = draw(st.integers(1, 4))
    widths = [draw(st.integers(0, 9)) for _ in range(n_node)]
    g, off = R.make_inventory(widths, mig=True)
    t = R.empty_table()
    for m in range(2):                                   # two random models: sizes 1..8, starts inside 8 slices
        for p in range(draw(st.integers(1, 6))):
            size = draw(st.sampled_from([1, 2, 3, 4, 7, 8]))
            starts = draw(st.lists(st.integers(0, 8 - size), min_size=0, max_size=4, unique=True))
            t[m, p] = (size, 0, R.mask_of(starts))
    for i in range(len(g)):
        g["model"][i] = draw(st.integers(0, 1))
        g["flags"][i] = draw(st.sampled_from([0, 1, 1, 1, 3, 5, 4, 2]))
        g["busy"][i] = draw(st.integers(0, 255)) if (g["flags"][i] & 1) and draw(st.booleans()) else 0
        g["mem_free_mib"][i] = draw(st.sampled_from([0, 1000, 16384]))
        g["share_cnt"][i] = draw(st.sampled_from([0, 0, 0, 2]))
    n_claim = draw(st.integers(0, 24))
    c = np.zeros(n_claim, dtype=R.CLAIM_DTYPE)
    gid = 1
    i = 0
    while i < n_claim:
        kind = draw(st.sampled_from([0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 1, 7]))
        node = n_node if draw(st.integers(0, 11)) == 0 else draw(st.integers(0, n_node - 1))   # n_node: names no node -> INVALID
        run = draw(st.integers(2, 4)) if kind == 1 and draw(st.integers(0, 4)) == 0 else 1
        for k in range(min(run, n_claim - i)):
            c[i]["kind"] = kind
            c[i]["node"] = node
            c[i]["profile"] = draw(st.sampled_from([0, 0, 1, 1, 2, 3, 4, 5, 6, 7, 16])) if kind == 1 else draw(st.sampled_from([0, 16, 200]))
            c[i]["count"] = draw(st.sampled_from([1, 1, 1, 1, 1, 2, 2, 3, 0, 33])) if kind == 0 else 1
            c[i]["mem_limit_mib"] = draw(st.sampled_from([0, 500, 1000, 20000])) if kind == 2 else 0
            c[i]["group"] = gid if run > 1 else 0
            i += 1
        gid += run > 1
    out_off, n_out = R.out_offsets(c, n_node)
    return g, off, t, c, out_off, n_out


@settings(max_examples=300, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)   # stable in CI; hammered with 10k random examples when written
@given(problems())
def test_oracle_agrees_with_naive_and_keeps_the_spec_invariants(pkg, oracle, prob):
    from oracle import naive
    R = pkg.records
    g, off, t, c, out_off, n_out = prob
    out, after = oracle.allocate(g, off, t, c, out_off, n_out)
    tbl = [[(int(t[m, p]["size"]), int(t[m, p]["start_mask"])) for p in range(16)] for m in range(16)]
    nout, nafter = naive.allocate(_dicts(g), [int(x) for x in off], tbl, _dicts(c), [int(x) for x in out_off])
    assert [tuple(int(v) for v in r) for r in out][:len(nout)] == nout
    for a, b in zip(after, nafter):
        assert (int(a["busy"]), int(a["flags"]), int(a["mem_free_mib"]), int(a["share_cnt"])) == \
               (b["busy"], b["flags"], b["mem_free_mib"], b["share_cnt"])
    # invariants
    busy = g["busy"].astype(np.int64).copy()
    n_node = len(off) - 1
    slots = R.claim_slots(c, n_node)
    for i in range(len(c)):
        for k in range(int(slots[i])):
            o = out[int(out_off[i]) + k]
            if o["status"] != R.ST_OK:
                assert o["gpu"] == R.GPU_NONE
                continue
            gi = int(o["gpu"])
            assert off[c[i]["node"]] <= gi < off[c[i]["node"] + 1]           # on the selected node
            if c[i]["kind"] == R.KIND_MIG:
                e = t[g["model"][gi], c[i]["profile"]]
                assert o["size"] == e["size"] and (int(e["start_mask"]) >> int(o["start"])) & 1   # a table placement
                m = ((1 << int(o["size"])) - 1) << int(o["start"])
                assert busy[gi] & m == 0                                          # never overlaps (memory slices)
                busy[gi] |= m
                assert g["flags"][gi] & R.GPU_MIG_ENABLED
            elif c[i]["kind"] == R.KIND_GPU:
                assert not (g["flags"][gi] & R.GPU_MIG_ENABLED)                   # GPU xor MIG parent (nvlib.go:152)
    assert (after["busy"].astype(np.int64) == busy).all()
    # UnsuitableNodes (spec §8) from its definition, with the independent restatement: a pod is suitable on a node iff
    # allocating just its claims there, on a snapshot, leaves every slot OK
    if len(c):
        cut = [i for i in range(1, len(c)) if c["group"][i] == 0 or c["group"][i] != c["group"][i - 1]]
        pod_off = np.array([0] + cut[::2] + [len(c)], dtype=np.uint32)         # pods never split a co-location run
        n_pod = len(pod_off) - 1
        cand_nodes = np.tile(np.arange(n_node, dtype=np.uint32), n_pod)
        cand_off = (np.arange(n_pod + 1, dtype=np.uint32) * n_node).astype(np.uint32)
        bits = oracle.unsuitable(g, off, t, c, pod_off, cand_nodes, cand_off)
        got = np.unpackbits(bits, bitorder="little")[: n_pod * n_node].reshape(n_pod, n_node)
        gd, offl = _dicts(g), [int(x) for x in off]
        for pd in range(n_pod):
            pc = c[pod_off[pd]: pod_off[pd + 1]].copy()
            for nd in range(n_node):
                pc["node"] = nd
                po, pn = R.out_offsets(pc, n_node)
                res, _ = naive.allocate(gd, offl, tbl, _dicts(pc), [int(x) for x in po])
                assert bool(got[pd, nd]) == all(r[4] == R.ST_OK for r in res[:pn]), (pd, nd)
    # Deallocate everything that was allocated: back to the start (spec §9)
    back = oracle.deallocate(after, c, out, out_off, n_node=n_node)
    assert back.tobytes() == g.tobytes()
