"""Per-kernel device times (event pairs around every kernel, dra_set_profiling) of the sort path on the larger configs:
  python profiles/stage_times.py  ->  profiles/r02_stage_times.txt"""
import importlib, os, sys, numpy as np, torch, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("k8s-dra-driver_b200")
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ctx = pkg.api.Context(device=0, stream=s.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, mk in (("cfg3", pkg.synth.cfg3), ("cfg5", pkg.synth.cfg5), ("1M x 80k", lambda: pkg.synth.cfg2(1_000_000, 10_000))):
    w = mk()
    ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
    d_claims = torch.from_numpy(w.claims.view(np.uint8).copy()).cuda()
    d_out = torch.zeros(w.n_out * 8, dtype=torch.uint8, device="cuda")
    step = lambda: ctx.allocate_device(d_claims.data_ptr(), w.n_claim, None, d_out.data_ptr(), w.n_out, pkg.api.F_FRESH_INVENTORY)
    ts = []
    for it in range(25):
        flush.fill_(1); e0.record(s); step(); e1.record(s); ctx.sync()
        if it >= 5: ts.append(e0.elapsed_time(e1) * 1e3)
    ctx.set_profiling(True); acc = {}
    for it in range(25):
        flush.fill_(1); step(); ctx.sync()
        if it >= 5:
            for k, v in ctx.timings_us().items(): acc.setdefault(k, []).append(v)
    ctx.set_profiling(False)
    st = {k: round(statistics.median(v), 1) for k, v in acc.items() if statistics.median(v) > 0}
    print(f"{name:<10} {w.n_claim:>8} claims {w.n_node:>6} nodes: step {statistics.median(ts):7.1f} us (PDL chain) | kernels alone (serialised, event pairs): {st}  sum {sum(st.values()):.1f}")
ctx.close()
