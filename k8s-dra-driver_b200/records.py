"""Flat record layouts of the allocation path (spec/ALLOCATION.md §1) as numpy dtypes.

These mirror ``include/dra_alloc.h`` byte for byte.  The reference types they flatten:
``GpuInfo`` / ``MigDeviceInfo`` / ``MigProfileInfo`` / ``MigDevicePlacement``
(cmd/nvidia-dra-plugin/deviceinfo.go:30-64), NVML ``GpuInstancePlacement{Start,Size}``
(vendor/github.com/NVIDIA/go-nvml/pkg/nvml/types_gen.go:755-758) and the k8s
``DeviceRequestAllocationResult`` (vendor/k8s.io/api/resource/v1beta1/types.go:795-840).
"""
from __future__ import annotations

import numpy as np

MAX_GPUS_PER_NODE = 32
MAX_MODELS = 16
MAX_PROFILES = 16
MAX_COUNT = 32
MAX_GROUP = 32
GPU_NONE = 0xFFFFFFFF

KIND_GPU, KIND_MIG, KIND_SHARED = 0, 1, 2
GPU_MIG_ENABLED, GPU_FULL_ALLOCATED, GPU_UNAVAILABLE = 0x01, 0x02, 0x04
PROFILE_GPU, PROFILE_SHARED = 0xFF, 0xFE
ST_OK, ST_NO_CAPACITY, ST_BAD_PROFILE, ST_GROUP, ST_MEM_LIMIT, ST_INVALID = 0, 1, 2, 3, 4, 5
ST_POD, ST_SEARCH_LIMIT = 6, 7            # pod mode (spec §12)
MAX_POD, EXH_BUDGET = 32, 4096
STATUS_NAMES = {0: "OK", 1: "NO_CAPACITY", 2: "BAD_PROFILE", 3: "GROUP", 4: "MEM_LIMIT", 5: "INVALID", 6: "POD",
                7: "SEARCH_LIMIT"}

F_NODE_SORTED = 0x1
F_EXHAUSTIVE = 0x4

GPU_DTYPE = np.dtype([("busy", "<u2"), ("flags", "u1"), ("model", "u1"), ("mem_free_mib", "<u4"),
                      ("node", "<u4"), ("share_cnt", "<u2"), ("rsvd", "<u2")])
CLAIM_DTYPE = np.dtype([("kind", "u1"), ("profile", "u1"), ("count", "<u2"), ("node", "<u4"),
                        ("mem_limit_mib", "<u4"), ("group", "<u4")])
OUT_DTYPE = np.dtype([("gpu", "<u4"), ("start", "u1"), ("size", "u1"), ("profile", "u1"),
                      ("status", "u1")])
PROF_DTYPE = np.dtype([("size", "u1"), ("rsvd", "u1"), ("start_mask", "<u2")])

# selectors (spec §10)
ATTR_DTYPE = np.dtype([("mem_total_mib", "<u4"), ("cc", "<u4"), ("index", "<u4"), ("product", "<u2"),
                       ("driver_major", "<u2")])
SEL_INS_DTYPE = np.dtype([("op", "u1"), ("attr", "u1"), ("cmp", "u1"), ("rsvd", "u1"), ("value", "<u4")])
SEL_MAX_INS = 8
SEL_END, SEL_CMP, SEL_AND, SEL_OR, SEL_NOT = 0, 1, 2, 3, 4
ATTR_MEMORY_MIB, ATTR_CC, ATTR_INDEX, ATTR_PRODUCT, ATTR_DRIVER_MAJOR = 0, 1, 2, 3, 4
CMP_EQ, CMP_NE, CMP_LT, CMP_LE, CMP_GT, CMP_GE, CMP_IN_MASK = 0, 1, 2, 3, 4, 5, 6
assert ATTR_DTYPE.itemsize == 16 and SEL_INS_DTYPE.itemsize == 8


def selector(*ins) -> np.ndarray:
    """Build one selector program from ('cmp', attr, cmp, value) / 'and' / 'or' / 'not' items (postfix)."""
    prog = np.zeros(SEL_MAX_INS, dtype=SEL_INS_DTYPE)
    assert len(ins) <= SEL_MAX_INS
    for i, it in enumerate(ins):
        if it == "and":
            prog[i]["op"] = SEL_AND
        elif it == "or":
            prog[i]["op"] = SEL_OR
        elif it == "not":
            prog[i]["op"] = SEL_NOT
        else:
            _, attr, cmp_, val = it
            prog[i] = (SEL_CMP, attr, cmp_, 0, val)
    return prog


assert GPU_DTYPE.itemsize == 16 and CLAIM_DTYPE.itemsize == 16
assert OUT_DTYPE.itemsize == 8 and PROF_DTYPE.itemsize == 4

# NVML GI profile enum (vendor/github.com/NVIDIA/go-nvml/pkg/nvml/const.go:745-766)
GI_1_SLICE, GI_2_SLICE, GI_3_SLICE, GI_4_SLICE, GI_7_SLICE = 0, 1, 2, 3, 4
GI_8_SLICE, GI_6_SLICE, GI_1_SLICE_REV1, GI_2_SLICE_REV1, GI_1_SLICE_REV2 = 5, 6, 7, 8, 9
GI_PROFILE_COUNT = 10

# compute slices per GI enum: go-nvlib NewMigProfile, mig_profile.go:57-84
GI_COMPUTE_SLICES = {0: 1, 7: 1, 9: 1, 1: 2, 8: 2, 2: 3, 3: 4, 6: 6, 4: 7, 5: 8}


def empty_table() -> np.ndarray:
    """Placement table: [MAX_MODELS, MAX_PROFILES] of ProfEnt, all "not offered"."""
    return np.zeros((MAX_MODELS, MAX_PROFILES), dtype=PROF_DTYPE)


def mask_of(starts) -> int:
    m = 0
    for s in starts:
        m |= 1 << int(s)
    return m


def a100_40gb_rows() -> dict:
    """SYNTHETIC model-0 table (spec appendix; external knowledge, not in /root/reference):
    A100-40GB geometry in memory-slice units.  enum -> (size, starts)."""
    return {
        GI_1_SLICE: (1, range(0, 7)),
        GI_2_SLICE: (2, (0, 2, 4)),
        GI_3_SLICE: (4, (0, 4)),
        GI_4_SLICE: (4, (0,)),
        GI_7_SLICE: (8, (0,)),
        GI_1_SLICE_REV1: (1, range(0, 7)),
        GI_1_SLICE_REV2: (2, (0, 2, 4, 6)),
    }


# a100-40gb profile names, for the codec (<N>g.<M>gb form of go-nvlib mig_profile.go:145-154)
A100_40GB_NAMES = {GI_1_SLICE: "1g.5gb", GI_2_SLICE: "2g.10gb", GI_3_SLICE: "3g.20gb",
                   GI_4_SLICE: "4g.20gb", GI_7_SLICE: "7g.40gb", GI_1_SLICE_REV1: "1g.5gb+me",
                   GI_1_SLICE_REV2: "1g.10gb"}


def default_table() -> np.ndarray:
    """Model 0 = synthetic A100-40GB; model 1 = a synthetic 'half' part (4 slices) used by the
    heterogeneous tests: 1g size 1 starts 0..3, 2g size 2 starts 0,2, 4g size 4 start 0."""
    t = empty_table()
    for p, (size, starts) in a100_40gb_rows().items():
        t[0, p] = (size, 0, mask_of(starts))
    t[1, GI_1_SLICE] = (1, 0, mask_of(range(4)))
    t[1, GI_2_SLICE] = (2, 0, mask_of((0, 2)))
    t[1, GI_4_SLICE] = (4, 0, mask_of((0,)))
    return t


def make_inventory(gpus_per_node, *, mig=True, model=0, mem_free_mib=40960, busy=0):
    """Convenience: homogeneous nodes.  gpus_per_node: list of GPU counts per node."""
    n_gpu = int(sum(gpus_per_node))
    g = np.zeros(n_gpu, dtype=GPU_DTYPE)
    node_off = np.zeros(len(gpus_per_node) + 1, dtype=np.uint32)
    node_off[1:] = np.cumsum(gpus_per_node)
    g["node"] = np.repeat(np.arange(len(gpus_per_node), dtype=np.uint32), gpus_per_node)
    g["flags"] = GPU_MIG_ENABLED if mig else 0
    g["model"] = model
    g["mem_free_mib"] = mem_free_mib
    g["busy"] = busy
    return g, node_off


def claim_slots(claims: np.ndarray, n_node: int, have_off: bool = True) -> np.ndarray:
    """slots(c) of spec §1/§3."""
    cnt = claims["count"].astype(np.int64)
    ok = (claims["kind"] == KIND_GPU) & (claims["node"] < n_node) & (cnt >= 1) & (cnt <= MAX_COUNT)
    if not have_off:
        ok &= cnt == 1
    return np.where(ok, cnt, 1).astype(np.uint32)


def out_offsets(claims: np.ndarray, n_node: int):
    """Exclusive prefix of slots: (out_off[n_claim] uint32, n_out)."""
    s = claim_slots(claims, n_node, True)
    off = np.zeros(len(claims), dtype=np.uint32)
    if len(claims):
        off[1:] = np.cumsum(s[:-1])
    return off, int(s.sum())
