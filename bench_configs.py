#!/usr/bin/env python
"""Extra measurement (not the driver's bench): every BASELINE.json config on one B200, device-resident and
end-to-end, next to the CPU oracle (1 thread), with the parity check inline.  One JSON line per config."""
import importlib, json, os, statistics, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("k8s-dra-driver_b200")
from oracle import oracle as O
R, S = pkg.records, pkg.synth
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ctx = pkg.api.Context(device=0, stream=s.cuda_stream, flags=pkg.api.CFG_NO_FUSED if os.environ.get("DRA_BENCH_NO_FUSED") else 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
F = pkg.api.F_FRESH_INVENTORY
for name in ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]:
    w = S.CONFIGS[name]()
    ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
    ref, _ = O.allocate(w.gpus, w.node_off, w.table, w.claims, w.out_off, w.n_out)
    d_claims = torch.from_numpy(w.claims.view(np.uint8).copy()).cuda()
    d_out = torch.zeros(w.n_out * 8, dtype=torch.uint8, device="cuda")
    l0 = ctx.launch_count()
    ctx.allocate_device(d_claims.data_ptr(), w.n_claim, None, d_out.data_ptr(), w.n_out, F); ctx.sync()
    launches = ctx.launch_count() - l0
    assert d_out.cpu().numpy().tobytes() == ref.tobytes(), name
    ts = []
    for it in range(60):
        flush.fill_(1)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(s); ctx.allocate_device(d_claims.data_ptr(), w.n_claim, None, d_out.data_ptr(), w.n_out, F); b.record(s)
        torch.cuda.synchronize()
        if it >= 10: ts.append(a.elapsed_time(b) * 1e3)
    dev_us = statistics.median(ts)
    pc = pkg.api.PinnedBuffer(w.n_claim, R.CLAIM_DTYPE); pc.array[:] = w.claims
    po = pkg.api.PinnedBuffer(w.n_out, R.OUT_DTYPE)
    es = []
    for it in range(40):
        t0 = time.perf_counter(); ctx.allocate(pc.array, None, w.n_out, flags=F, out=po.array); es.append(time.perf_counter() - t0)
    e2e_us = statistics.median(es[5:]) * 1e6
    assert po.array.tobytes() == ref.tobytes()
    cs = []
    for it in range(7):
        t0 = time.perf_counter(); O.allocate(w.gpus, w.node_off, w.table, w.claims, w.out_off, w.n_out); cs.append(time.perf_counter() - t0)
    cpu_us = statistics.median(cs) * 1e6
    print(json.dumps({"config": name, "claims": w.n_claim, "gpus": w.n_gpu, "nodes": w.n_node, "kernel_launches": launches,
                      "device_us": round(dev_us, 2), "device_alloc_per_s": round(w.n_claim / dev_us * 1e6),
                      "e2e_us": round(e2e_us, 2), "e2e_alloc_per_s": round(w.n_claim / e2e_us * 1e6),
                      "cpu_oracle_1t_us": round(cpu_us, 1), "cpu_alloc_per_s": round(w.n_claim / cpu_us * 1e6),
                      "allocated": int((ref["status"] == 0).sum()), "algorithmic_bytes": w.algorithmic_bytes(),
                      "hbm_frac": round(w.algorithmic_bytes() / (dev_us * 1e-6) / 6587.7e9, 6), "parity": "bit-exact"}))
    pc.free(); po.free()
