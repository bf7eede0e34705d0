"""Pod mode (spec §12) timings on one GPU: allocate_pods / unsuitable with DRA_F_EXHAUSTIVE, repeated calls (the bench line's
pod_mode keys are single shots), with the per-kernel event times of the last call."""
import importlib, os, sys, time, statistics
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("k8s-dra-driver_b200")
A, S = pkg.api, pkg.synth
pw, ppo = S.pods(20_000, 64, 1)
sub = 4000
po = ppo[: sub + 1]; pc = pw.claims[: po[-1]]
with A.Context(device=0) as ctx:
    ctx.set_table(pw.table); ctx.set_inventory(pw.gpus, pw.node_off)
    for name, fn in (("allocate_pods exhaustive, 20k pods / 60k claims, 64 nodes", lambda: ctx.allocate_pods(pw.claims, ppo, flags=A.F_EXHAUSTIVE | A.F_FRESH_INVENTORY)),
                     ("allocate_pods first-fit (same pods)", lambda: ctx.allocate_pods(pw.claims, ppo, flags=A.F_FRESH_INVENTORY)),
                     ("unsuitable exhaustive, 4000 pods x 64 nodes", lambda: ctx.unsuitable(pc, po, flags=A.F_EXHAUSTIVE)),
                     ("unsuitable first-fit, 4000 pods x 64 nodes", lambda: ctx.unsuitable(pc, po))):
        ctx.set_inventory(pw.gpus, pw.node_off)       # every measurement against the same (pre-fragmented, otherwise free) inventory
        ts = []
        for _ in range(6):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        ctx.set_inventory(pw.gpus, pw.node_off)
        ctx.set_profiling(True); fn(); k = {a: round(b, 1) for a, b in ctx.timings_us().items() if b > 0}; ctx.set_profiling(False)
        print(f"{name:<62} first call {ts[0] * 1e3:7.2f} ms, then median {statistics.median(ts[1:]) * 1e3:7.2f} ms; kernels (us): {k}")
