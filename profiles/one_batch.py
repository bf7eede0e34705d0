"""One device-resident cfg2 Allocate batch (plus warm-up), for `ncu --set full -k regex:k_fused -s 3 -c 1`."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("k8s-dra-driver_b200")
w = pkg.synth.cfg2()
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ctx = pkg.api.Context(device=0, stream=s.cuda_stream)
ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
d_claims = torch.from_numpy(w.claims.view(np.uint8).copy()).cuda()
d_out = torch.zeros(w.n_out * 8, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(5):
    flush.fill_(1)
    ctx.allocate_device(d_claims.data_ptr(), w.n_claim, None, d_out.data_ptr(), w.n_out, pkg.api.F_FRESH_INVENTORY)
    ctx.sync()
print("ok")
