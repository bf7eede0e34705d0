"""Generates tests/golden/alloc_*.npz: inputs + the CPU oracle's outputs for small allocation cases.

The reference has no implementation or fixtures for this path, so these are the
oracle's answers, frozen: the CPU suite checks that the oracle still reproduces them (drift), the GPU
suite checks that the CUDA path reproduces them without needing the oracle on the box.  The in-tree
workload shape they start from is demo/specs/quickstart/gpu-test4.yaml:19-44 with the geometry of
demo/specs/quickstart/mig-parted-config.yaml:8-16 (GPUs 0-3 MIG-enabled, 4-7 not).
Run from the repo root:  python tests/golden/make_alloc_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
P = importlib.import_module("k8s-dra-driver_b200")
R, S = P.records, P.synth
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def gpu_test4(replicas=4):
    """4 replicas of one pod whose claim has 4 MIG requests that must share a parent GPU."""
    g, off = R.make_inventory([8], mig=True)
    g["flags"][4:] = 0                      # half-balanced: devices 4..7 mig-enabled: false
    c = np.zeros(4 * replicas, dtype=R.CLAIM_DTYPE)
    c["kind"] = R.KIND_MIG
    c["count"] = 1
    c["profile"] = [R.GI_1_SLICE, R.GI_1_SLICE, R.GI_2_SLICE, R.GI_3_SLICE] * replicas
    c["group"] = np.repeat(np.arange(1, replicas + 1, dtype=np.uint32), 4)
    w = S.Workload("gpu_test4", g, off, R.default_table(), c)
    return w.finish()


def cases():
    yield gpu_test4()
    yield S.cfg1()
    w = S.cfg2(600, 7); w.name = "cfg2_small"; yield w
    w = S.cfg4(900, 3); w.name = "cfg4_small"; yield w
    w = S.cfg5(700, 5); w.name = "cfg5_small"; yield w
    for seed in range(4):
        yield S.mixed(800, 13, seed)
    w = S.mixed(300, 5, 11, invalid=False); w.name = "mixed11_valid"; yield w
    # homogeneous all-MIG nodes: the packers' fast loops (<= 8 and > 8 GPUs per node, 8- and 16-slice parts)
    w = S.homog(900, 6, 0, max_width=8); w.name = "homog_w8"; yield w
    w = S.homog(1500, 7, 1, max_width=32); w.name = "homog_w32"; yield w
    w = S.homog(1200, 6, 4, wide16=True, max_width=32); w.name = "homog_16slice"; yield w


def b200_table(n_claim=420, n_node=8, seed=3):
    """Every MIG profile a real B200 offers (tests/golden/b200_gi_profiles.json: NVML's answers, recorded on the GPU box), on
    nodes of 8 such GPUs, a quarter of them with slices 2-3 taken."""
    import json
    doc = json.load(open(os.path.join(HERE, "b200_gi_profiles.json")))
    t = R.default_table()
    for p, (size, mask) in enumerate(doc["row"]):
        t[15, p] = (size, 0, mask)
    g, off = R.make_inventory([8] * n_node, model=15, mem_free_mib=183359)
    r = S.splitmix64(0xB200 + seed, 3 * n_claim)
    offered = np.array([p for p, (size, _) in enumerate(doc["row"]) if size], dtype=np.uint8)
    c = np.zeros(n_claim, dtype=R.CLAIM_DTYPE)
    c["kind"], c["count"] = R.KIND_MIG, 1
    c["profile"] = offered[(r[0::3] % np.uint64(len(offered))).astype(np.int64)]
    c["node"] = (r[1::3] % np.uint64(n_node)).astype(np.uint32)
    g["busy"][(r[2::3][: len(g)] % np.uint64(4) == 0)] = 0b1100
    return S.Workload("b200_table", g, off, t, c).finish()


def main():
    only = set(sys.argv[1:])                 # e.g. `make_alloc_golden.py b200_table`: (re)generate just these cases
    for w in list(cases()) + [b200_table()]:
        if only and w.name not in only:
            continue
        out, after = O.allocate(w.gpus, w.node_off, w.table, w.claims, w.out_off, w.n_out)
        path = os.path.join(HERE, f"alloc_{w.name}.npz")
        np.savez_compressed(path, gpus=w.gpus, node_off=w.node_off, table=w.table, claims=w.claims,
                            out_off=(w.out_off if w.out_off is not None else np.zeros(0, np.uint32)),
                            out=out, gpus_after=after)
        print(f"{path}: {len(w.claims)} claims, status histogram {np.bincount(out['status'], minlength=6)}")


if __name__ == "__main__":
    main()
