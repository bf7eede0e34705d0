"""Where does the single-launch kernel (k_fused) stop paying against the sort path (hist/scan/scatter/pack)?
cfg2-like batches of growing size, device-resident, both paths.   python profiles/path_crossover.py  (GPU box)"""
import importlib, os, statistics, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("k8s-dra-driver_b200")
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
F = pkg.api.F_FRESH_INVENTORY
ctxs = {"fused": pkg.api.Context(device=0, stream=s.cuda_stream), "sort": pkg.api.Context(device=0, stream=s.cuda_stream, flags=pkg.api.CFG_NO_FUSED)}
print("n_node n_claim   fused_us  sort_us  launches(fused)")
for n_node in (64, 125, 250):
    for n_claim in (6000, 10000, 11000, 12000, 16000, 20000, 24000, 32000, 48000):
        w = pkg.synth.cfg2(n_claim, n_node)
        d_claims = torch.from_numpy(w.claims.view(np.uint8).copy()).cuda()
        d_out = torch.zeros(w.n_out * 8, dtype=torch.uint8, device="cuda")
        row = {}
        for name, ctx in ctxs.items():
            ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
            l0 = ctx.launch_count()
            ctx.allocate_device(d_claims.data_ptr(), w.n_claim, None, d_out.data_ptr(), w.n_out, F); ctx.sync()
            row[name + "_l"] = ctx.launch_count() - l0
            ts = []
            for it in range(25):
                flush.fill_(1)
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(s); ctx.allocate_device(d_claims.data_ptr(), w.n_claim, None, d_out.data_ptr(), w.n_out, F); b.record(s)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            row[name] = statistics.median(ts[5:])
        print(f"{n_node:6d} {n_claim:7d}   {row['fused']:8.2f} {row['sort']:8.2f}   {row['fused_l']}")
