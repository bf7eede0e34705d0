"""CPU suite: property-based check of the multi-GPU host logic (shard.py).  For arbitrary problems and
world sizes 1..5: plan() covers the nodes with contiguous disjoint ranges; every claim is owned by exactly one rank;
allocating every rank's local batch (oracle) and merging equals allocating the global batch — OutRecs in input order
with global GPU indices, and the per-rank inventories concatenate to the global inventory-after."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from test_oracle_properties import problems


@settings(max_examples=200, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(problems(), st.integers(1, 5))
def test_sharded_allocation_equals_global(pkg, oracle, prob, world):
    g, off, t, c, out_off, n_out = prob
    n_node = len(off) - 1
    ranges = pkg.shard.plan(c["node"], n_node, world)
    assert ranges[0][0] == 0 and ranges[-1][1] == n_node
    assert all(a1 == b0 and a0 <= a1 for (a0, a1), (b0, _) in zip(ranges, ranges[1:]))
    ref_out, ref_after = oracle.allocate(g, off, t, c, out_off, n_out)
    parts, afters, owned = [], [], np.zeros(len(c), dtype=np.int64)
    for r in range(world):
        lb = pkg.shard.local_batch(g, off, c, r, ranges)
        owned[lb.sel] += 1
        lout, lafter = oracle.allocate(lb.gpus, lb.node_off, t, lb.claims, lb.out_off, lb.n_out)
        parts.append((lb.sel, lb.out_off, lout, lb.gpu_base))
        la = lafter.copy(); la["node"] += np.uint32(lb.n0)
        afters.append(la)
    assert (owned == 1).all()                              # every claim on exactly one rank (strays on rank 0)
    merged = pkg.shard.merge(n_out, out_off, parts)
    assert merged.tobytes() == ref_out.tobytes()
    assert np.concatenate(afters).tobytes() == ref_after.tobytes()
