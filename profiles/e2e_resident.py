#!/usr/bin/env python
"""Where the end-to-end call goes in resident mode: host wall clock of dra_allocate_batch (bare C-ABI call
through ctypes, pinned buffers) split into  ctypes overhead | doorbell -> seen | seen -> egress issued | fences | completion
word -> host, next to the cooperative-launch path and the copy-engine path.  cfg2, one B200."""
import importlib, os, statistics, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("k8s-dra-driver_b200")
R, A = pkg.records, pkg.api
w = pkg.synth.cfg2()
pc = A.PinnedBuffer(w.n_claim, R.CLAIM_DTYPE); pc.array[:] = w.claims
po = A.PinnedBuffer(w.n_out, R.OUT_DTYPE)
F = A.F_FRESH_INVENTORY


def timed(ctx, n=400):
    for _ in range(20):
        ctx.allocate_raw(pc.ptr, w.n_claim, None, po.ptr, w.n_out, F)
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); ctx.allocate_raw(pc.ptr, w.n_claim, None, po.ptr, w.n_out, F); ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e6, min(ts) * 1e6


for name, flags in (("resident kernel + doorbell", A.CFG_RESIDENT), ("one cooperative launch per batch (direct host I/O)", 0),
                    ("copy engine H2D -> kernel -> D2H as one CUDA graph", A.CFG_NO_DIRECT | A.CFG_USE_GRAPH)):
    with A.Context(device=0, flags=flags) as ctx:
        ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
        med, best = timed(ctx)
        line = f"{name:<58} median {med:6.2f} us   min {best:6.2f} us"
        if flags == A.CFG_RESIDENT:
            t = ctx.debug_serve_times().astype(np.int64)
            line += (f"\n    on the GPU (globaltimer): doorbell seen -> CTA 0's egress issued {(t[1] - t[0]) / 1e3:5.2f} us, -> every CTA fenced "
                     f"{(t[2] - t[0]) / 1e3:5.2f} us, -> completion word out {(t[3] - t[0]) / 1e3:5.2f} us;  the rest of the {med:.1f} us is the doorbell's "
                     f"way to the GPU (PCIe read poll), the completion word's way back, and the caller")
        print(line)
    ts = []
lib = A.load()
for _ in range(2000):
    t0 = time.perf_counter(); lib.dra_abi_version(); ts.append(time.perf_counter() - t0)
print(f"ctypes call of an empty C function: median {statistics.median(ts) * 1e6:.2f} us (the binding's own cost; a cgo call is of the same order)")
