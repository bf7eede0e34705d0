"""One device-resident batch of 1M claims on 10k nodes x 8 GPUs (sort path with wave-sized hist tiles and the row-wise scan),
three times — the target of the ncu capture of the large-batch kernels."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("k8s-dra-driver_b200")
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ctx = pkg.api.Context(device=0, stream=s.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
w = pkg.synth.cfg2(1_000_000, 10_000)
ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
d_claims = torch.from_numpy(w.claims.view(np.uint8).copy()).cuda()
d_out = torch.zeros(w.n_out * 8, dtype=torch.uint8, device="cuda")
for it in range(3):
    flush.fill_(1)
    ctx.allocate_device(d_claims.data_ptr(), w.n_claim, None, d_out.data_ptr(), w.n_out, pkg.api.F_FRESH_INVENTORY)
    ctx.sync()
print("ok")
