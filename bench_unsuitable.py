#!/usr/bin/env python
"""Extra measurement (not the driver's bench): UnsuitableNodes on the cfg2 inventory.

10,000 pods (one cfg2 claim each; every 8th pod has a 3-request co-located claim) x 125 candidate nodes
= 1.25 M (pod, node) evaluations per call against a half-full inventory, CUDA path through the C ABI vs the
CPU oracle.  Prints one JSON line."""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("k8s-dra-driver_b200")
from oracle import oracle as O
R, S = pkg.records, pkg.synth

w = S.cfg2()
_, inv = O.allocate(w.gpus, w.node_off, w.table, w.claims[:4000])          # half-full inventory
n_pod = 10_000
claims = w.claims[:n_pod].copy()
pod_off = np.arange(n_pod + 1, dtype=np.uint32)
cand_nodes = np.tile(np.arange(w.n_node, dtype=np.uint32), n_pod)
cand_off = (np.arange(n_pod + 1, dtype=np.uint32) * w.n_node).astype(np.uint32)
n_pair = int(cand_off[-1])

with pkg.api.Context(device=0) as ctx:
    ctx.set_table(w.table); ctx.set_inventory(inv, w.node_off)
    bits_sparse = ctx.unsuitable(claims, pod_off, cand_nodes, cand_off)       # explicit candidate lists
    bits = ctx.unsuitable(claims, pod_off)                                     # dense form: every node, no lists
    assert bits.tobytes() == bits_sparse.tobytes()
    ts, tsp = [], []
    for _ in range(20):
        t0 = time.perf_counter(); ctx.unsuitable(claims, pod_off); ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); ctx.unsuitable(claims, pod_off, cand_nodes, cand_off); tsp.append(time.perf_counter() - t0)
    ctx.set_profiling(True); ctx.unsuitable(claims, pod_off)
    k_us = list(ctx.timings_us().values())[0]
t0 = time.perf_counter(); ref = O.unsuitable(inv, w.node_off, w.table, claims[:1000], pod_off[:1001], cand_nodes[:125000], cand_off[:1001]); cpu_s = time.perf_counter() - t0
assert bits[: 125000 // 8].tobytes() == ref[: 125000 // 8].tobytes()
e2e = float(np.median(ts))
print(json.dumps({"metric": "UnsuitableNodes (pod, node) evaluations/s", "pairs": n_pair, "suitable_pairs": int(np.unpackbits(bits).sum()),
                  "e2e_ms": e2e * 1e3, "e2e_pairs_per_s": n_pair / e2e, "e2e_ms_with_candidate_lists": float(np.median(tsp)) * 1e3, "kernel_us": k_us, "kernel_pairs_per_s": n_pair / (k_us * 1e-6),
                  "cpu_oracle_pairs_per_s": 125000 / cpu_s, "cpu_sample": "first 1000 pods x 125 nodes, 1 thread", "parity": "bit-exact on the sample"}))
