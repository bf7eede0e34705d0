"""ctypes binding of libdra_alloc.so (include/dra_alloc.h) — the same entry points a Go driver binds
with cgo.  This module is plumbing: every call below crosses the C ABI and runs the
sm_100a kernels; nothing is computed in Python and nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import records as R

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libdra_alloc.so")

OK, E_INVAL, E_CUDA, E_NCCL, E_NOMEM, E_STATE = 0, -1, -2, -3, -4, -5
CFG_USE_GRAPH, CFG_NO_FUSED, CFG_NO_DIRECT, CFG_RESIDENT = 0x1, 0x2, 0x4, 0x8
F_NODE_SORTED, F_FRESH_INVENTORY, F_EXHAUSTIVE = 0x1, 0x2, 0x4


class DraError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libdra_alloc: rc={code}: {msg}")
        self.code = code


class _Cfg(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("stream", C.c_void_p),
                ("max_claims", C.c_uint32), ("flags", C.c_uint32)]


# every symbol include/dra_alloc.h declares: name -> (restype, argtypes)
_vp, _u32, _i32, _u64 = C.c_void_p, C.c_uint32, C.c_int, C.c_uint64
SYMBOLS = {
    "dra_abi_version": (_i32, []),
    "dra_ctx_create": (_i32, [C.POINTER(_Cfg), C.POINTER(_vp)]),
    "dra_ctx_destroy": (None, [_vp]),
    "dra_last_error": (C.c_char_p, [_vp]),
    "dra_set_placement_table": (_i32, [_vp, _u32, _vp]),
    "dra_set_inventory": (_i32, [_vp, _vp, _u32, _vp, _u32]),
    "dra_set_gpu_attrs": (_i32, [_vp, _vp, _u32]),
    "dra_set_selectors": (_i32, [_vp, _vp, _u32]),
    "dra_get_inventory": (_i32, [_vp, _vp, _u32]),
    "dra_reset_inventory": (_i32, [_vp]),
    "dra_allocate_batch": (_i32, [_vp, _vp, _u32, _vp, _vp, _u32, _u32]),
    "dra_allocate_batch_device": (_i32, [_vp, _vp, _u32, _vp, _vp, _u32, _u32]),
    "dra_ctx_sync": (_i32, [_vp]),
    "dra_unsuitable_batch": (_i32, [_vp, _vp, _u32, _vp, _u32, _vp, _vp, _vp, _u32]),
    "dra_allocate_pods_batch": (_i32, [_vp, _vp, _u32, _vp, _u32, _vp, _vp, _u32, _u32]),
    "dra_deallocate_batch": (_i32, [_vp, _vp, _u32, _vp, _vp, _u32]),
    "dra_comm_unique_id": (_i32, [_vp]),
    "dra_comm_init": (_i32, [_vp, _vp, _i32, _i32]),
    "dra_allocate_batch_gather_device": (_i32, [_vp, _vp, _u32, _vp, _vp, _u32, _u32, _u32]),
    "dra_gather_table": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_u32)]),
    "dra_gather_read": (_i32, [_vp, _vp, _u32]),
    "dra_peer_export": (_i32, [_vp, _u32, _vp]),
    "dra_peer_import": (_i32, [_vp, _vp]),
    "dra_comm_init_local": (_i32, [_vp, _i32, _i32]),
    "dra_peer_import_local": (_i32, [_vp, _vp]),
    "dra_peer_rendezvous_device": (_i32, [_vp]),
    "dra_set_shard": (_i32, [_vp, _u32, _u32, _i32]),
    "dra_set_shard_map": (_i32, [_vp, _vp, _i32]),
    "dra_shard_export": (_i32, [_vp, _u32, _u32, _vp]),
    "dra_allocate_batch_global_device": (_i32, [_vp, _vp, _u32, _vp, _u32, _u32]),
    "dra_mps_limits_batch": (_i32, [_vp, _vp, _u32, _vp, _vp]),
    "dra_imex_offsets_batch": (_i32, [_vp, _vp, _vp, _u32, C.c_int32, C.c_int32, _vp]),
    "dra_calibrate": (_i32, [_vp, C.POINTER(_u32)]),
    "dra_serve_start": (_i32, [_vp]),
    "dra_serve_stop": (_i32, [_vp]),
    "dra_serve_batches": (_u64, [_vp]),
    "dra_host_alloc": (_vp, [C.c_size_t]),
    "dra_host_free": (None, [_vp]),
    "dra_launch_count": (_u64, [_vp]),
    "dra_set_profiling": (_i32, [_vp, _i32]),
    "dra_debug_timeline": (_i32, [_vp, _vp, _u32]),
    "dra_debug_shard_times": (_i32, [_vp, _vp]),
    "dra_debug_serve_times": (_i32, [_vp, _vp]),
    "dra_debug_noop": (_i32, [_vp, _u32, _u32, _u32]),
    "dra_get_timings": (_i32, [_vp, _vp, _i32]),
}

_lib = None


def load(build_if_stale: bool = True) -> C.CDLL:
    """dlopen the in-tree library; raises if it is missing — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_stale:
        from . import build as _b
        try:
            if _b.stale():
                _b.build()
        except RuntimeError:
            if not os.path.exists(SO_PATH):
                raise
    if not os.path.exists(SO_PATH):
        raise FileNotFoundError(f"{SO_PATH} not built: run `python __graft_entry__.py build`; "
                                "the allocation path has no CPU fallback")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the ABI and the header drift apart
        fn.restype, fn.argtypes = res, args
    if lib.dra_abi_version() != 2:
        raise RuntimeError("libdra_alloc ABI version mismatch")
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


class PinnedBuffer:
    """Page-locked host array from dra_host_alloc: passing it to Context calls skips staging."""

    def __init__(self, n: int, dtype):
        self._lib = load()
        dtype = np.dtype(dtype)
        self.nbytes = max(16, int(n) * dtype.itemsize)
        self.ptr = self._lib.dra_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError("dra_host_alloc failed")
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(n))

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.dra_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One dra_ctx: one CUDA device, one stream, one inventory."""

    def __init__(self, device: int = 0, stream: int | None = None, max_claims: int = 0, flags: int = 0):
        self._lib = load()
        cfg = _Cfg(2, device, C.c_void_p(stream) if stream else None, max_claims, flags)
        h = C.c_void_p()
        rc = self._lib.dra_ctx_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise DraError(rc, (self._lib.dra_last_error(None) or b"").decode())
        self._h = h
        self.n_gpu = 0
        self.n_node = 0

    # -- plumbing -----------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != OK:
            raise DraError(rc, (self._lib.dra_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dra_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state ----------------------------------------------------------------------------------------
    def set_table(self, table: np.ndarray):
        """table: PROF_DTYPE[MAX_MODELS, MAX_PROFILES]"""
        t = np.ascontiguousarray(table, dtype=R.PROF_DTYPE)
        assert t.shape == (R.MAX_MODELS, R.MAX_PROFILES)
        for m in range(R.MAX_MODELS):
            self._check(self._lib.dra_set_placement_table(self._h, m, _ptr(t[m])))

    def set_inventory(self, gpus: np.ndarray, node_off: np.ndarray):
        g = np.ascontiguousarray(gpus, dtype=R.GPU_DTYPE)
        off = np.ascontiguousarray(node_off, dtype=np.uint32)
        self._check(self._lib.dra_set_inventory(self._h, _ptr(g), len(g), _ptr(off), len(off) - 1))
        self.n_gpu, self.n_node = len(g), len(off) - 1

    def set_gpu_attrs(self, attrs: np.ndarray | None):
        if attrs is None or len(attrs) == 0:
            self._check(self._lib.dra_set_gpu_attrs(self._h, None, 0))
            return
        a = np.ascontiguousarray(attrs, dtype=R.ATTR_DTYPE)
        self._check(self._lib.dra_set_gpu_attrs(self._h, _ptr(a), len(a)))

    def set_selectors(self, sels: np.ndarray | None):
        if sels is None or len(sels) == 0:
            self._check(self._lib.dra_set_selectors(self._h, None, 0))
            return
        s_ = np.ascontiguousarray(sels, dtype=R.SEL_INS_DTYPE).reshape(-1, R.SEL_MAX_INS)
        self._check(self._lib.dra_set_selectors(self._h, _ptr(s_), len(s_)))

    def get_inventory(self) -> np.ndarray:
        g = np.zeros(self.n_gpu, dtype=R.GPU_DTYPE)
        self._check(self._lib.dra_get_inventory(self._h, _ptr(g), self.n_gpu))
        return g

    def reset_inventory(self):
        self._check(self._lib.dra_reset_inventory(self._h))

    def sync(self):
        self._check(self._lib.dra_ctx_sync(self._h))

    # -- the path ---------------------------------------------------------------------------------------
    def allocate(self, claims: np.ndarray, out_off: np.ndarray | None = None, n_out: int | None = None,
                 flags: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        c = claims if (claims.dtype == R.CLAIM_DTYPE and claims.flags.c_contiguous) else \
            np.ascontiguousarray(claims, dtype=R.CLAIM_DTYPE)
        oo = None if out_off is None else np.ascontiguousarray(out_off, dtype=np.uint32)
        if n_out is None:
            n_out = len(c) if oo is None else int(R.claim_slots(c, self.n_node).sum())
        if out is None:
            out = np.zeros(n_out, dtype=R.OUT_DTYPE)
        self._check(self._lib.dra_allocate_batch(self._h, _ptr(c), len(c), _ptr(oo), _ptr(out), n_out, flags))
        return out

    def allocate_raw(self, claims_ptr: int, n_claim: int, out_off_ptr: int | None, out_ptr: int, n_out: int,
                     flags: int = 0) -> None:
        """dra_allocate_batch on raw HOST addresses (e.g. PinnedBuffer.ptr): the bare C-ABI call, without the
        numpy conveniences of allocate() — what a cgo caller pays."""
        rc = self._lib.dra_allocate_batch(self._h, claims_ptr, n_claim, out_off_ptr, out_ptr, n_out, flags)
        if rc != OK:
            self._check(rc)

    def allocate_device(self, d_claims: int, n_claim: int, d_out_off: int | None, d_out: int, n_out: int,
                        flags: int = 0):
        """Device pointers (ints); enqueues on the context's stream, no synchronisation."""
        self._check(self._lib.dra_allocate_batch_device(self._h, _ptr(d_claims), n_claim, _ptr(d_out_off),
                                                        _ptr(d_out), n_out, flags))

    def allocate_gather_device(self, d_claims: int, n_claim: int, d_out_off: int | None, d_out_all: int,
                               n_out: int, n_per_rank: int, flags: int = 0):
        self._check(self._lib.dra_allocate_batch_gather_device(self._h, _ptr(d_claims), n_claim,
                                                               _ptr(d_out_off), _ptr(d_out_all), n_out,
                                                               n_per_rank, flags))

    def allocate_pods(self, claims, pod_off, out_off=None, n_out=None, flags: int = 0) -> np.ndarray:
        """Pod mode (spec §12): every pod atomically; flags & F_EXHAUSTIVE = backtracking placement search."""
        c = np.ascontiguousarray(claims, dtype=R.CLAIM_DTYPE)
        po = np.ascontiguousarray(pod_off, dtype=np.uint32)
        oo = None if out_off is None else np.ascontiguousarray(out_off, dtype=np.uint32)
        if n_out is None:
            n_out = len(c) if oo is None else int(R.claim_slots(c, self.n_node).sum())
        out = np.zeros(n_out, dtype=R.OUT_DTYPE)
        self._check(self._lib.dra_allocate_pods_batch(self._h, _ptr(c), len(c), _ptr(po), len(po) - 1, _ptr(oo), _ptr(out),
                                                      n_out, flags))
        return out

    def unsuitable(self, claims, pod_off, cand_nodes=None, cand_off=None, flags: int = 0) -> np.ndarray:
        """cand_nodes/cand_off None = dense form: every pod against every node (bit k = pod * n_node + node).
        flags & F_EXHAUSTIVE: spec §12 — unsuitable only if NO assignment of the pod exists."""
        c = np.ascontiguousarray(claims, dtype=R.CLAIM_DTYPE)
        po = np.ascontiguousarray(pod_off, dtype=np.uint32)
        if cand_nodes is None:
            n_pair = (len(po) - 1) * self.n_node
            bits = np.zeros((n_pair + 7) // 8 + 8, dtype=np.uint8)
            self._check(self._lib.dra_unsuitable_batch(self._h, _ptr(c), len(c), _ptr(po), len(po) - 1, None, None, _ptr(bits), flags))
            return bits[: (n_pair + 7) // 8]
        cn = np.ascontiguousarray(cand_nodes, dtype=np.uint32)
        co = np.ascontiguousarray(cand_off, dtype=np.uint32)
        n_pair = int(co[-1])
        bits = np.zeros((n_pair + 7) // 8 + 8, dtype=np.uint8)
        self._check(self._lib.dra_unsuitable_batch(self._h, _ptr(c), len(c), _ptr(po), len(po) - 1,
                                                   _ptr(cn), _ptr(co), _ptr(bits), flags))
        return bits[: (n_pair + 7) // 8]

    def deallocate(self, claims, out, out_off=None):
        c = np.ascontiguousarray(claims, dtype=R.CLAIM_DTYPE)
        o = np.ascontiguousarray(out, dtype=R.OUT_DTYPE)
        oo = None if out_off is None else np.ascontiguousarray(out_off, dtype=np.uint32)
        self._check(self._lib.dra_deallocate_batch(self._h, _ptr(c), len(c), _ptr(oo), _ptr(o), len(o)))

    # -- adjacent integer searches ------------------------------------------------------------------------
    def mps_limits(self, nbytes) -> tuple:
        b = np.ascontiguousarray(nbytes, dtype=np.int64)
        mib = np.zeros(len(b), dtype=np.int64); valid = np.zeros(len(b), dtype=np.uint8)
        self._check(self._lib.dra_mps_limits_batch(self._h, _ptr(b), len(b), _ptr(mib), _ptr(valid)))
        return mib, valid.astype(bool)

    def imex_offsets(self, used_lists, step: int = 128, limit: int = 2048) -> np.ndarray:
        off = np.zeros(len(used_lists) + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(u) for u in used_lists])
        used = np.ascontiguousarray(np.concatenate([np.asarray(u, dtype=np.int32) for u in used_lists]) if len(used_lists) and off[-1] else np.zeros(1, np.int32), dtype=np.int32)
        out = np.zeros(len(used_lists), dtype=np.int32)
        self._check(self._lib.dra_imex_offsets_batch(self._h, _ptr(used), _ptr(off), len(used_lists), step, limit, _ptr(out)))
        return out

    # -- multi-GPU ------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = load().dra_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc != OK:
            raise DraError(rc, (load().dra_last_error(None) or b"").decode())
        return bytes(buf)

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._check(self._lib.dra_comm_init(self._h, C.cast(buf, C.c_void_p), rank, world))

    def peer_rendezvous(self) -> None:
        """Enqueue a device-side rendezvous of all ranks on this context's stream (dra_peer_rendezvous_device)."""
        self._check(self._lib.dra_peer_rendezvous_device(self._h))

    def gather_read(self, out_all: np.ndarray) -> np.ndarray:
        """D2H of the last gather's table into out_all (OUT_DTYPE, world * n_per_rank); synchronises."""
        self._check(self._lib.dra_gather_read(self._h, _ptr(out_all), len(out_all)))
        return out_all

    def peer_disable(self) -> None:
        self._check(self._lib.dra_peer_export(self._h, 0, None))

    def peer_export(self, n_per_rank: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self._lib.dra_peer_export(self._h, n_per_rank, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def comm_init_local(self, rank: int, world: int) -> None:
        """Rank / world of a context whose peers live in THIS process (no NCCL communicator)."""
        self._check(self._lib.dra_comm_init_local(self._h, rank, world))

    def peer_import_local(self, ctxs) -> None:
        """Map the gather buffers of the world contexts of this process (rank order)."""
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        self._check(self._lib.dra_peer_import_local(self._h, C.cast(arr, C.c_void_p)))

    def set_shard(self, node_lo: int, node_hi: int, take_stray: bool = False) -> None:
        self._check(self._lib.dra_set_shard(self._h, node_lo, node_hi, 1 if take_stray else 0))

    def set_shard_map(self, bounds, stray_rank: int = 0) -> None:
        """bounds: world + 1 node indices; this rank serves [bounds[rank], bounds[rank + 1])."""
        b = np.ascontiguousarray(bounds, dtype=np.uint32)
        self._check(self._lib.dra_set_shard_map(self._h, _ptr(b), stray_rank))

    def shard_export(self, n_out_max: int, cap_per_rank: int = 0, want_handle: bool = True) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self._lib.dra_shard_export(self._h, n_out_max, cap_per_rank, C.cast(buf, C.c_void_p) if want_handle else None))
        return bytes(buf)

    def allocate_global_device(self, d_claims: int, n_claim: int, d_out_off: int | None, n_out: int, flags: int = 0):
        """Sharded global batch: the SAME claim array on every rank (device pointer); enqueues, no synchronisation.
        The result table: gather_read()."""
        self._check(self._lib.dra_allocate_batch_global_device(self._h, _ptr(d_claims), n_claim, _ptr(d_out_off), n_out, flags))

    def peer_import(self, handles) -> None:
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._check(self._lib.dra_peer_import(self._h, C.cast(buf, C.c_void_p)))

    def calibrate(self) -> int:
        """Measure the single-launch / sort-path crossover on this device + inventory; returns the claim count."""
        v = _u32(0)
        self._check(self._lib.dra_calibrate(self._h, C.byref(v)))
        return int(v.value)

    # -- resident mode ----------------------------------------------------------------------------------
    def serve_start(self) -> None:
        self._check(self._lib.dra_serve_start(self._h))

    def serve_stop(self) -> None:
        self._check(self._lib.dra_serve_stop(self._h))

    def serve_batches(self) -> int:
        return int(self._lib.dra_serve_batches(self._h))

    # -- instrumentation ------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self._lib.dra_launch_count(self._h))

    def set_profiling(self, on: bool):
        self._check(self._lib.dra_set_profiling(self._h, 1 if on else 0))

    def debug_noop(self, grid: int, block: int, smem: int) -> None:
        self._check(self._lib.dra_debug_noop(self._h, grid, block, smem))

    def debug_timeline(self, n_cta: int) -> np.ndarray:
        buf = np.zeros(n_cta * 8, dtype=np.uint64)
        n = self._lib.dra_debug_timeline(self._h, _ptr(buf), len(buf))
        return buf[:n].reshape(-1, 8)

    def debug_serve_times(self):
        buf = np.zeros(4, dtype=np.uint64)
        return buf if self._lib.dra_debug_serve_times(self._h, _ptr(buf)) == 4 else None

    def debug_shard_times(self):
        buf = np.zeros(2, dtype=np.uint64)
        return buf if self._lib.dra_debug_shard_times(self._h, _ptr(buf)) == 2 else None

    def timings_us(self) -> dict:
        buf = (C.c_float * 5)()
        n = self._lib.dra_get_timings(self._h, C.cast(buf, C.c_void_p), 5)
        names = ["bucket_hist", "bucket_scan", "bucket_scatter", "pack", "all_gather"]
        return {names[i]: float(buf[i]) for i in range(n)}
