"""Multi-GPU host logic: shard a batch by NODE across ranks and merge the gathered results.

Allocation state is per node (spec §2; the reference's world view is per node too — one plugin, one pool
per node, vendor/k8s.io/dynamic-resource-allocation/kubeletplugin/draplugin.go:427-435), so nodes are
sharded whole and no rank ever needs another rank's inventory.  The only exchange is the all-gather of the
OutRecs.  Everything here is index arithmetic on the host; the allocation itself runs in
libdra_alloc.so on each rank's GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import records as R


def plan(claim_nodes: np.ndarray, n_node: int, world: int) -> list[tuple[int, int]]:
    """Contiguous node ranges [n0, n1), one per rank, balanced by the number of claims that select them."""
    per_node = np.bincount(claim_nodes[claim_nodes < n_node].astype(np.int64), minlength=n_node)
    cum = np.concatenate([[0], np.cumsum(per_node)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        n = int(np.searchsorted(cum, target, side="left"))
        bounds.append(min(max(n, bounds[-1]), n_node))
    bounds.append(n_node)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


@dataclass
class LocalBatch:
    rank: int
    n0: int
    n1: int
    gpu_base: int                 # global index of the rank's first GPU
    gpus: np.ndarray              # GPU_DTYPE, node renumbered from 0
    node_off: np.ndarray
    claims: np.ndarray            # CLAIM_DTYPE, node renumbered; input order preserved
    sel: np.ndarray               # global claim index of each local claim
    out_off: np.ndarray           # local first slot of each local claim
    n_out: int


def local_batch(gpus, node_off, claims, rank: int, ranges) -> LocalBatch:
    """The part of a global batch that rank `rank` owns.  Claims naming no node go to rank 0 (they come
    back INVALID, spec §3)."""
    n_node = len(node_off) - 1
    n0, n1 = ranges[rank]
    g0, g1 = int(node_off[n0]), int(node_off[n1])
    lg = np.ascontiguousarray(gpus[g0:g1]).copy()
    lg["node"] -= np.uint32(n0)
    loff = (node_off[n0:n1 + 1] - np.uint32(g0)).astype(np.uint32)
    mine = (claims["node"] >= n0) & (claims["node"] < n1)
    if rank == 0:
        mine |= claims["node"] >= n_node
    sel = np.nonzero(mine)[0].astype(np.uint32)
    lc = np.ascontiguousarray(claims[sel]).copy()
    stray = lc["node"] >= n_node
    lc["node"] = np.where(stray, np.uint32(0xFFFFFFFF), lc["node"] - np.uint32(n0))
    # slots are a property of the claim and of "does it name a node at all"
    slots = R.claim_slots(claims[sel], n_node, True)
    off = np.zeros(len(sel), dtype=np.uint32)
    if len(sel):
        off[1:] = np.cumsum(slots[:-1])
    return LocalBatch(rank, n0, n1, g0, lg, loff, lc, sel, off, int(slots.sum()))


def merge(n_out_global: int, global_out_off: np.ndarray, parts) -> np.ndarray:
    """parts: iterable of (sel, local_out_off, local_out, gpu_base) from every rank (after the all-gather).
    Returns the OutRecs in global input order with global GPU indices."""
    out = np.zeros(n_out_global, dtype=R.OUT_DTYPE)
    for sel, loff, lout, gpu_base in parts:
        lout = lout.copy()
        ok = lout["gpu"] != R.GPU_NONE
        lout["gpu"][ok] += np.uint32(gpu_base)
        if len(sel) == 0:
            continue
        ends = np.concatenate([loff[1:], [len(lout)]]).astype(np.int64)
        lens = ends - loff.astype(np.int64)
        src = np.arange(len(lout), dtype=np.int64)
        dst = np.repeat(global_out_off[sel].astype(np.int64) - loff.astype(np.int64), lens) + src
        out[dst] = lout
    return out
