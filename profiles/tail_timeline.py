#!/usr/bin/env python
"""Where the multi-GPU step goes: per-CTA globaltimer stamps of the gather tail of k_fused.

  DRA_TIMELINE=1 python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/tail_timeline.py

One global batch of N x 10k claims per step (bench.py's workload).  Per rank, relative to the first CTA's entry:
  entry (last CTA in)  |  pack done (max over CTAs)  |  packets sent  |  header/ticket  |  all peers' records in
plus the event-timed step and the compaction kernel alone."""
import importlib, os, statistics, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("k8s-dra-driver_b200")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
w = pkg.synth.cfg2(10_000 * world, 125 * world, 8)
ctx = pkg.api.Context(device=local, stream=stream.cuda_stream)
ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
uid = [pkg.api.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(uid[0], rank, world)
ranges = pkg.shard.plan(w.claims["node"], w.n_node, world)
ctx.set_shard_map([r[0] for r in ranges] + [ranges[-1][1]], stray_rank=0)
hs = [None] * world
dist.all_gather_object(hs, ctx.shard_export(w.n_out)); ctx.peer_import(hs)
d_claims = torch.from_numpy(w.claims.view(np.uint8).copy()).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
F = pkg.api.F_FRESH_INVENTORY
n_local = ranges[rank][1] - ranges[rank][0]
rows, steps = [], []
for it in range(40):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    a.record(stream); ctx.allocate_global_device(d_claims.data_ptr(), w.n_claim, None, w.n_out, F); b.record(stream)
    ctx.sync()
    if it < 8:
        continue
    steps.append(a.elapsed_time(b) * 1e3)
    tl = ctx.debug_timeline(2 * n_local + 6)
    t = tl[n_local + 4: n_local + 4 + n_local + 1].astype(np.int64)       # tail stamps: [entry, pack done, sent, ticket, received]
    e0 = t[:, 0].min()
    sc = ctx.debug_shard_times()
    comp = (float(int(sc[1]) - int(sc[0])) / 1e3, float(int(e0) - int(sc[1])) / 1e3) if sc is not None else (0.0, 0.0)
    rows.append([comp[0], comp[1], (t[:, 0].max() - e0) / 1e3, (t[:, 1].max() - e0) / 1e3, (t[:, 2].max() - e0) / 1e3, (t[:, 3].max() - e0) / 1e3, (t[:, 4].max() - e0) / 1e3])
med = [statistics.median(r[k] for r in rows) for k in range(7)]
out = [None] * world
dist.all_gather_object(out, (rank, statistics.median(steps), med))
if rank == 0:
    print(f"world {world}: one global batch of {w.n_claim} claims, {n_local} nodes per rank; medians over {len(rows)} steps, us")
    print("rank  step(events)  | compaction  gap to k_fused | last CTA in  pack done  packets sent  ticket/header  all records in   (from k_fused's first CTA, globaltimer)")
    for r, st, m in sorted(out):
        print(f"{r:4d}  {st:11.2f}   | {m[0]:10.2f} {m[1]:14.2f} | {m[2]:10.2f} {m[3]:10.2f} {m[4]:13.2f} {m[5]:14.2f} {m[6]:15.2f}")
ctx.close()
dist.destroy_process_group()
