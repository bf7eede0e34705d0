"""ctypes wrapper of oracle/_build/libdra_oracle.so — the CPU oracle of spec/ALLOCATION.md.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package never imports this module.
PARITY UNPINNED for the claim search (no reference implementation exists in the snapshot);
pinned against the reference's own vectors where they exist (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdra_oracle.so")

GPU_DTYPE = np.dtype([("busy", "<u2"), ("flags", "u1"), ("model", "u1"), ("mem_free_mib", "<u4"),
                      ("node", "<u4"), ("share_cnt", "<u2"), ("rsvd", "<u2")])
CLAIM_DTYPE = np.dtype([("kind", "u1"), ("profile", "u1"), ("count", "<u2"), ("node", "<u4"),
                        ("mem_limit_mib", "<u4"), ("group", "<u4")])
OUT_DTYPE = np.dtype([("gpu", "<u4"), ("start", "u1"), ("size", "u1"), ("profile", "u1"),
                      ("status", "u1")])


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("dra_oracle.c", "dra_oracle.h")]
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(s) > os.path.getmtime(_SO) for s in src if os.path.exists(s)):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
        _lib.dra_oracle_allocate.argtypes = [vp, u32, vp, u32, vp, vp, u32, vp, vp, u32]
        _lib.dra_oracle_allocate.restype = i32
        _lib.dra_oracle_allocate_mt.argtypes = [vp, u32, vp, u32, vp, vp, u32, vp, vp, u32, i32]
        _lib.dra_oracle_allocate_mt.restype = i32
        _lib.dra_oracle_unsuitable.argtypes = [vp, u32, vp, u32, vp, vp, u32, vp, u32, vp, vp, vp]
        _lib.dra_oracle_unsuitable.restype = i32
        _lib.dra_oracle_allocate_pods.argtypes = [vp, u32, vp, u32, vp, vp, u32, vp, u32, vp, vp, u32, u32]
        _lib.dra_oracle_allocate_pods.restype = i32
        _lib.dra_oracle_unsuitable_ex.argtypes = [vp, u32, vp, u32, vp, vp, u32, vp, u32, vp, vp, vp, u32]
        _lib.dra_oracle_unsuitable_ex.restype = i32
        _lib.dra_oracle_deallocate.argtypes = [vp, u32, u32, vp, u32, vp, vp, u32]
        _lib.dra_oracle_deallocate.restype = i32
        _lib.dra_oracle_set_selectors.argtypes = [vp, u32, vp, u32]
        _lib.dra_oracle_set_selectors.restype = None
        _lib.dra_oracle_megabyte.argtypes = [C.c_int64, C.POINTER(C.c_int)]
        _lib.dra_oracle_megabyte.restype = C.c_int64
        _lib.dra_oracle_imex_offset.argtypes = [vp, u32, C.c_int32, C.c_int32]
        _lib.dra_oracle_imex_offset.restype = C.c_int32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


def allocate(gpus, node_off, table, claims, out_off=None, n_out=None, threads: int = 1):
    """Returns (out, gpus_after).  Inputs are not modified."""
    g = _c(gpus, GPU_DTYPE).copy()
    off = _c(node_off, np.uint32)
    t = np.ascontiguousarray(table)
    c = _c(claims, CLAIM_DTYPE)
    oo = None if out_off is None else _c(out_off, np.uint32)
    if n_out is None:
        n_out = len(c) if oo is None else 0
    out = np.zeros(max(n_out, 1), dtype=OUT_DTYPE)
    rc = lib().dra_oracle_allocate_mt(_p(g), len(g), _p(off), len(off) - 1, _p(t), _p(c), len(c),
                                      _p(oo), _p(out), n_out, threads)
    if rc != 0:
        raise ValueError(f"dra_oracle_allocate: rc={rc}")
    return out[:n_out], g


F_EXHAUSTIVE = 0x4


def allocate_pods(gpus, node_off, table, claims, pod_off, out_off=None, n_out=None, flags: int = 0):
    """spec §12 (pod mode).  Returns (out, gpus_after)."""
    g = _c(gpus, GPU_DTYPE).copy()
    off = _c(node_off, np.uint32)
    t = np.ascontiguousarray(table)
    c = _c(claims, CLAIM_DTYPE)
    po = _c(pod_off, np.uint32)
    oo = None if out_off is None else _c(out_off, np.uint32)
    if n_out is None:
        n_out = len(c) if oo is None else 0
    out = np.zeros(max(n_out, 1), dtype=OUT_DTYPE)
    rc = lib().dra_oracle_allocate_pods(_p(g), len(g), _p(off), len(off) - 1, _p(t), _p(c), len(c), _p(po), len(po) - 1,
                                        _p(oo), _p(out), n_out, flags)
    if rc != 0:
        raise ValueError(f"dra_oracle_allocate_pods: rc={rc}")
    return out[:n_out], g


def unsuitable(gpus, node_off, table, claims, pod_off, cand_nodes, cand_off, flags: int = 0):
    if flags:
        return _unsuitable_ex(gpus, node_off, table, claims, pod_off, cand_nodes, cand_off, flags)
    g = _c(gpus, GPU_DTYPE)
    off = _c(node_off, np.uint32)
    t = np.ascontiguousarray(table)
    c = _c(claims, CLAIM_DTYPE)
    po = _c(pod_off, np.uint32)
    cn = _c(cand_nodes, np.uint32)
    co = _c(cand_off, np.uint32)
    n_pair = int(co[-1])
    bits = np.zeros((n_pair + 7) // 8 + 1, dtype=np.uint8)
    rc = lib().dra_oracle_unsuitable(_p(g), len(g), _p(off), len(off) - 1, _p(t), _p(c), len(c),
                                     _p(po), len(po) - 1, _p(cn), _p(co), _p(bits))
    if rc != 0:
        raise ValueError(f"dra_oracle_unsuitable: rc={rc}")
    return bits[: (n_pair + 7) // 8]


def _unsuitable_ex(gpus, node_off, table, claims, pod_off, cand_nodes, cand_off, flags):
    g = _c(gpus, GPU_DTYPE)
    off = _c(node_off, np.uint32)
    t = np.ascontiguousarray(table)
    c = _c(claims, CLAIM_DTYPE)
    po = _c(pod_off, np.uint32)
    cn = _c(cand_nodes, np.uint32)
    co = _c(cand_off, np.uint32)
    n_pair = int(co[-1])
    bits = np.zeros((n_pair + 7) // 8 + 1, dtype=np.uint8)
    rc = lib().dra_oracle_unsuitable_ex(_p(g), len(g), _p(off), len(off) - 1, _p(t), _p(c), len(c),
                                        _p(po), len(po) - 1, _p(cn), _p(co), _p(bits), flags)
    if rc != 0:
        raise ValueError(f"dra_oracle_unsuitable_ex: rc={rc}")
    return bits[: (n_pair + 7) // 8]


def deallocate(gpus, claims, out, out_off=None, n_node=None):
    """n_node: number of nodes of the inventory (claims naming no node have one slot, spec §9); None = every node
    index is taken as valid."""
    g = _c(gpus, GPU_DTYPE).copy()
    c = _c(claims, CLAIM_DTYPE)
    o = _c(out, OUT_DTYPE)
    oo = None if out_off is None else _c(out_off, np.uint32)
    rc = lib().dra_oracle_deallocate(_p(g), len(g), 0xFFFFFFFF if n_node is None else int(n_node), _p(c), len(c),
                                     _p(oo), _p(o), len(o))
    if rc != 0:
        raise ValueError(f"dra_oracle_deallocate: rc={rc}")
    return g


_sel_keep = []


def set_selectors(attrs=None, sels=None):
    """Selector context for the following calls (spec §10); call with no arguments to clear."""
    _sel_keep.clear()
    if sels is None or len(sels) == 0:
        lib().dra_oracle_set_selectors(None, 0, None, 0)
        return
    a = np.ascontiguousarray(attrs) if attrs is not None else np.zeros(0, dtype=np.uint8)
    s_ = np.ascontiguousarray(sels).reshape(-1, 8)
    _sel_keep.extend([a, s_])                  # the C side keeps the pointers
    lib().dra_oracle_set_selectors(_p(a) if len(a) else None, len(a), _p(s_), len(s_))


def megabyte(nbytes: int):
    v = C.c_int(0)
    r = lib().dra_oracle_megabyte(int(nbytes), C.byref(v))
    return int(r), bool(v.value)


def imex_offset(used, step=128, limit=2048) -> int:
    u = np.ascontiguousarray(used, dtype=np.int32)
    return int(lib().dra_oracle_imex_offset(_p(u), len(u), step, limit))
