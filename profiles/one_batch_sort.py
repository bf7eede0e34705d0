"""One device-resident cfg3 Allocate batch on ONE GPU (sort path: k_bucket_hist8 -> k_bucket_scan8 -> k_bucket_scatter -> k_pack)
plus one UnsuitableNodes batch (k_unsuitable, default and exhaustive) plus one sharded call with world 1 (k_shard_compact_flat
+ k_fused reading a device-side count), for the ncu captures of profiles/final_run_r02.sh."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("k8s-dra-driver_b200")
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ctx = pkg.api.Context(device=0, stream=s.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
F = pkg.api.F_FRESH_INVENTORY
w = pkg.synth.cfg3()
ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
d_claims = torch.from_numpy(w.claims.view(np.uint8).copy()).cuda()
d_out = torch.zeros(w.n_out * 8, dtype=torch.uint8, device="cuda")
for it in range(2):
    flush.fill_(1)
    ctx.allocate_device(d_claims.data_ptr(), w.n_claim, None, d_out.data_ptr(), w.n_out, F)
    ctx.sync()
w2 = pkg.synth.cfg2()
ctx.set_inventory(w2.gpus, w2.node_off)
pod_off = np.arange(10_001, dtype=np.uint32)
for it in range(1):
    ctx.unsuitable(w2.claims, pod_off)
    ctx.unsuitable(w2.claims, pod_off, flags=pkg.api.F_EXHAUSTIVE)
ctx.set_shard(0, w2.n_node, True)
d2 = torch.from_numpy(w2.claims.view(np.uint8).copy()).cuda()
for it in range(2):
    flush.fill_(1)
    ctx.allocate_global_device(d2.data_ptr(), w2.n_claim, None, w2.n_out, F)
    ctx.sync()
print("ok")
