"""Placement tables from NVML — the enumeration of deviceLib.getGpuInfo (cmd/nvidia-dra-plugin/nvlib.go:244-295).

For every GPU-instance profile id 0..GPU_INSTANCE_PROFILE_COUNT-1 (go-nvml const.go:745-766):
  nvmlDeviceGetGpuInstanceProfileInfo -> NOT_SUPPORTED / INVALID_ARGUMENT: skip (nvlib.go:247-252)
  nvmlDeviceGetGpuInstancePossiblePlacements(profile.id) -> the {start, size} list (nvlib.go:257-274, go-nvml
  device.go:2117-2132), in MEMORY slices (nvml.h:9761-9765)
  then the (compute-instance profile j, engine profile k) loop keeps the profiles with G == C (nvlib.go:276-292) — one
  survivor per supported GI profile (two for the 1-slice ones: the CI enum has two 1-slice profiles), named <G>g.<GB>gb[+me] (go-nvlib mig_profile.go:57-154, :323-331).
`enumerate_profiles(nvml, device)` runs that loop against any object with the two NVML calls (the real library through
ctypes, or a fake in the CPU tests); `table_row` packs the result into a dra_profile_tbl row.  NVML is dlopen'ed by
soname; a box without MIG-capable GPUs simply yields an empty list.  Host-side only.
"""
from __future__ import annotations

import ctypes as C
import math

from . import records as R

GPU_INSTANCE_PROFILE_COUNT = 10                      # const.go:764-765
COMPUTE_INSTANCE_PROFILE_COUNT = 8                   # const.go: COMPUTE_INSTANCE_PROFILE_COUNT
COMPUTE_INSTANCE_ENGINE_PROFILE_COUNT = 1
NVML_SUCCESS, NVML_ERROR_INVALID_ARGUMENT, NVML_ERROR_NOT_SUPPORTED = 0, 2, 3
# compute slices of the CI profile enum (mig_profile.go:84-104): 1,2,3,4,7,8,6,1(rev1)
CI_SLICES = {0: 1, 1: 2, 2: 3, 3: 4, 4: 7, 5: 8, 6: 6, 7: 1}
MEDIA_EXT = {R.GI_1_SLICE_REV1, R.GI_2_SLICE_REV1}  # "+me" attribute (mig_profile.go:106-111)


def mig_memory_gb(total_device_memory: int, mig_memory_mb: int) -> int:
    """getMigMemorySizeGB (go-nvlib mig_profile.go:323-331), same float operations."""
    frac = (float(mig_memory_mb) * (1024 * 1024)) / float(total_device_memory)
    frac = math.ceil(frac * 8) / 8
    total_gb = float((total_device_memory + (1 << 30) - 1) // (1 << 30))
    return int(math.floor(frac * total_gb + 0.5))    # math.Round: half away from zero (positive here)


def profile_name(gi_enum: int, mem_mb: int, total_mem: int) -> str:
    g = R.GI_COMPUTE_SLICES[gi_enum]
    return f"{g}g.{mig_memory_gb(total_mem, mem_mb)}gb" + ("+me" if gi_enum in MEDIA_EXT else "")


def enumerate_profiles(nvml, device, total_memory_bytes: int):
    """nvml: object with gpu_instance_profile_info(device, i) -> (ret, info dict with id, memory_size_mb) and
    gpu_instance_possible_placements(device, profile_id) -> (ret, [(start, size), ...]).
    Returns [{enum, id, name, placements}] exactly as getGpuInfo would list them."""
    out = []
    for i in range(GPU_INSTANCE_PROFILE_COUNT):
        ret, info = nvml.gpu_instance_profile_info(device, i)
        if ret in (NVML_ERROR_NOT_SUPPORTED, NVML_ERROR_INVALID_ARGUMENT):
            continue
        if ret != NVML_SUCCESS:
            raise RuntimeError(f"error retrieving GpuInstanceProfileInfo for profile {i}: {ret}")
        ret, placements = nvml.gpu_instance_possible_placements(device, info["id"])
        if ret in (NVML_ERROR_NOT_SUPPORTED, NVML_ERROR_INVALID_ARGUMENT):
            continue
        if ret != NVML_SUCCESS:
            raise RuntimeError(f"error retrieving GpuInstancePossiblePlacements for profile {i}: {ret}")
        g = R.GI_COMPUTE_SLICES[i]
        for j in range(COMPUTE_INSTANCE_PROFILE_COUNT):
            for _k in range(COMPUTE_INSTANCE_ENGINE_PROFILE_COUNT):
                if CI_SLICES[j] != g:                 # nvlib.go:283: keep G == C only
                    continue
                out.append({"enum": i, "id": info["id"], "name": profile_name(i, info["memory_size_mb"], total_memory_bytes),
                            "placements": list(placements), "ci_profile": j})
    return out


def table_row(profiles) -> "R.np.ndarray":
    """One dra_profile_tbl row (PROF_DTYPE[16]) from enumerate_profiles' output."""
    row = R.np.zeros(R.MAX_PROFILES, dtype=R.PROF_DTYPE)
    for p in profiles:
        if not p["placements"]:
            continue
        size = p["placements"][0][1]
        if any(s != size for _, s in p["placements"]):
            raise ValueError(f"profile {p['name']}: placements differ in size")
        row[p["enum"]] = (size, 0, R.mask_of(st for st, _ in p["placements"]))
    return row


class _ProfileInfo(C.Structure):                      # nvmlGpuInstanceProfileInfo_t (go-nvml types_gen.go:760-772)
    _fields_ = [("id", C.c_uint), ("isP2pSupported", C.c_uint), ("sliceCount", C.c_uint), ("instanceCount", C.c_uint),
                ("multiprocessorCount", C.c_uint), ("copyEngineCount", C.c_uint), ("decoderCount", C.c_uint),
                ("encoderCount", C.c_uint), ("jpegCount", C.c_uint), ("ofaCount", C.c_uint), ("memorySizeMB", C.c_ulonglong)]


class _Placement(C.Structure):                        # nvmlGpuInstancePlacement_t (types_gen.go:755-758)
    _fields_ = [("start", C.c_uint), ("size", C.c_uint)]


class _Memory(C.Structure):
    _fields_ = [("total", C.c_ulonglong), ("free", C.c_ulonglong), ("used", C.c_ulonglong)]


class Nvml:
    """The real library (libnvidia-ml.so.1) behind the two calls enumerate_profiles needs."""

    def __init__(self):
        self.lib = C.CDLL("libnvidia-ml.so.1")
        rc = self.lib.nvmlInit_v2()
        if rc != NVML_SUCCESS:
            raise RuntimeError(f"nvmlInit_v2: {rc}")

    def close(self):
        self.lib.nvmlShutdown()

    def device(self, index: int):
        h = C.c_void_p()
        rc = self.lib.nvmlDeviceGetHandleByIndex_v2(C.c_uint(index), C.byref(h))
        if rc != NVML_SUCCESS:
            raise RuntimeError(f"nvmlDeviceGetHandleByIndex_v2({index}): {rc}")
        return h

    def memory_total(self, device) -> int:
        m = _Memory()
        rc = self.lib.nvmlDeviceGetMemoryInfo(device, C.byref(m))
        if rc != NVML_SUCCESS:
            raise RuntimeError(f"nvmlDeviceGetMemoryInfo: {rc}")
        return int(m.total)

    def mig_mode(self, device):
        cur, pend = C.c_uint(), C.c_uint()
        rc = self.lib.nvmlDeviceGetMigMode(device, C.byref(cur), C.byref(pend))
        return rc, int(cur.value), int(pend.value)

    def gpu_instance_profile_info(self, device, i: int):
        info = _ProfileInfo()
        rc = self.lib.nvmlDeviceGetGpuInstanceProfileInfo(device, C.c_uint(i), C.byref(info))
        return rc, {"id": int(info.id), "memory_size_mb": int(info.memorySizeMB), "slice_count": int(info.sliceCount),
                    "instance_count": int(info.instanceCount), "multiprocessor_count": int(info.multiprocessorCount),
                    "copy_engine_count": int(info.copyEngineCount)}

    def gpu_instance_possible_placements(self, device, profile_id: int):
        n = C.c_uint(0)
        fn = getattr(self.lib, "nvmlDeviceGetGpuInstancePossiblePlacements_v2", None) or self.lib.nvmlDeviceGetGpuInstancePossiblePlacements
        rc = fn(device, C.c_uint(profile_id), None, C.byref(n))
        if rc != NVML_SUCCESS or n.value == 0:
            return rc, []
        arr = (_Placement * n.value)()
        rc = fn(device, C.c_uint(profile_id), arr, C.byref(n))
        return rc, [(int(arr[k].start), int(arr[k].size)) for k in range(n.value)]


def table_from_nvml(index: int = 0):
    """(PROF_DTYPE[16] row, profiles list) of GPU `index` — what SetMigProfiles / dra_set_placement_table is fed with in
    a real deployment.  On a GPU where MIG queries are not supported the list is empty (every profile skipped)."""
    n = Nvml()
    try:
        dev = n.device(index)
        profs = enumerate_profiles(n, dev, n.memory_total(dev))
        return table_row(profs), profs
    finally:
        n.close()
