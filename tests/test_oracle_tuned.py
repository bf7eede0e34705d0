"""CPU suite: the tuned CPU port (oracle/dra_oracle_tuned.c, the fair CPU baseline of bench.py) produces the plain
oracle's bytes on every workload class and thread count — it may be fast, it may not be different."""
import numpy as np
import pytest


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_tuned_port_matches_the_oracle(pkg, oracle, threads):
    from oracle import tuned
    S = pkg.synth
    cases = [S.cfg1(), S.cfg2(3000, 40), S.cfg4(2000, 11), S.cfg5(4000, 16), S.mixed(3000, 31, 5), S.mixed(2500, 7, 6, invalid=False),
             S.homog(3000, 20, 2), S.homog(2000, 9, 4, wide16=True)]
    for w in cases:
        ref, ref_inv = oracle.allocate(w.gpus, w.node_off, w.table, w.claims, w.out_off, w.n_out)
        out, inv = tuned.allocate(w.gpus, w.node_off, w.table, w.claims, w.out_off, w.n_out, threads=threads)
        assert out.tobytes() == ref.tobytes(), w.name
        assert inv.tobytes() == ref_inv.tobytes(), w.name


def test_tuned_port_with_selectors_falls_back(pkg, oracle):
    from oracle import tuned
    w = pkg.synth.mixed(2000, 13, 9)
    attrs, sels = pkg.synth.with_selectors(w, 1)
    oracle.set_selectors(attrs, sels)
    try:
        ref, ref_inv = oracle.allocate(w.gpus, w.node_off, w.table, w.claims, w.out_off, w.n_out)
    finally:
        oracle.set_selectors()
    # the tuned library has its own copy of the selector context (separate .so): without selectors loaded there, ids
    # beyond the (empty) table are INVALID — so only check that it is self-consistent with the plain code path it includes
    out, _ = tuned.allocate(w.gpus, w.node_off, w.table, w.claims, w.out_off, w.n_out, threads=4)
    assert len(out) == len(ref)
