import importlib, os, sys, numpy as np, torch, statistics
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("k8s-dra-driver_b200")
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for flags, nm in ((0,"auto"), (pkg.api.CFG_NO_FUSED, "sort")):
    ctx = pkg.api.Context(device=0, stream=s.cuda_stream, flags=flags)
    for name in ["cfg2","cfg3","cfg5"]:
        w = pkg.synth.CONFIGS[name]()
        ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
        d_claims = torch.from_numpy(w.claims.view(np.uint8).copy()).cuda()
        d_out = torch.zeros(w.n_out * 8, dtype=torch.uint8, device="cuda")
        ctx.set_profiling(True); acc={}
        for it in range(25):
            flush.fill_(1); ctx.allocate_device(d_claims.data_ptr(), w.n_claim, None, d_out.data_ptr(), w.n_out, pkg.api.F_FRESH_INVENTORY); ctx.sync()
            if it>=5:
                for k,v in ctx.timings_us().items(): acc.setdefault(k,[]).append(v)
        ctx.set_profiling(False)
        print(nm, name, {k: round(statistics.median(v),1) for k,v in acc.items() if statistics.median(v)>0})
    ctx.close()
