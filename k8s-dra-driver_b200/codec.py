"""ResourceSlice devices  <->  flat records.

Reads the `resource.k8s.io/v1beta1` Device objects the reference publishes —
``GpuInfo.GetDevice`` (cmd/nvidia-dra-plugin/deviceinfo.go:98-142) and ``MigDeviceInfo.GetDevice``
(deviceinfo.go:144-206): attributes ``type``, ``index``, ``parentIndex``, ``profile``, capacity ``memory`` and one
``memorySlice<i>`` per occupied slice — into ``GpuRec[]`` + ``node_off``, and turns ``OutRec``s back into the
reference's device names (deviceinfo.go:74-80) / ``DeviceRequestAllocationResult`` dicts
(vendor/k8s.io/api/resource/v1beta1/types.go:795-840).  Host-side index arithmetic only.
"""
from __future__ import annotations

import re

import numpy as np

from . import records as R
from . import sharing

DRIVER_NAME = "gpu.nvidia.com"        # cmd/nvidia-dra-controller/imex.go:41
_SLICE = re.compile(r"^memorySlice(\d+)$")


def _attr(dev: dict, key: str, default=None):
    a = dev.get("basic", {}).get("attributes", {}).get(key)
    if a is None:
        return default
    for k in ("string", "int", "bool", "version"):
        if k in a:
            return a[k]
    return default


def canonical_name(index: int) -> str:
    """GpuInfo.CanonicalName, deviceinfo.go:74-76."""
    return f"gpu-{index}"


def canonical_mig_name(parent_index: int, gi_profile_id: int, start: int, size: int) -> str:
    """MigDeviceInfo.CanonicalName, deviceinfo.go:78-80."""
    return f"gpu-{parent_index}-mig-{gi_profile_id}-{start}-{size}"


_MIG_NAME = re.compile(r"^gpu-(\d+)-mig-(\d+)-(\d+)-(\d+)$")
_GPU_NAME = re.compile(r"^gpu-(\d+)$")


def parse_name(name: str):
    """Inverse of the two name forms: ('gpu', index) or ('mig', parent, profileId, start, size)."""
    m = _MIG_NAME.match(name)
    if m:
        return ("mig",) + tuple(int(x) for x in m.groups())
    m = _GPU_NAME.match(name)
    if m:
        return ("gpu", int(m.group(1)))
    raise ValueError(f"not a gpu.nvidia.com device name: {name!r}")


class Inventory:
    """Published devices of a set of nodes, flattened."""

    def __init__(self, pools: dict, model_of=lambda product: 0):
        """pools: {node name: [Device dict, ...]} (one ResourceSlice pool per node, draplugin.go:427-435)."""
        self.node_names = list(pools)
        rows, self.local_index = [], []
        node_off = [0]
        for n, name in enumerate(self.node_names):
            gpus: dict[int, dict] = {}
            for d in pools[name]:
                t = _attr(d, "type")
                if t == "gpu":                                   # published only when MIG is off, nvlib.go:152
                    i = int(_attr(d, "index"))
                    mem = d.get("basic", {}).get("capacity", {}).get("memory", {}).get("value", "0")
                    gpus[i] = dict(mig=False, busy=0, mem=sharing.quantity_value(mem) >> 20,
                                   model=model_of(_attr(d, "productName", "")))
                elif t == "mig":                                 # existing MIG devices of MIG-enabled parents
                    p = int(_attr(d, "parentIndex"))
                    g = gpus.setdefault(p, dict(mig=True, busy=0, mem=0, model=model_of(_attr(d, "productName", ""))))
                    g["mig"] = True
                    for cap in d.get("basic", {}).get("capacity", {}):
                        m = _SLICE.match(cap)                    # memorySlice<i>, deviceinfo.go:199-204
                        if m:
                            g["busy"] |= 1 << int(m.group(1))
            for i in sorted(gpus):
                g = gpus[i]
                rows.append((g["busy"], R.GPU_MIG_ENABLED if g["mig"] else 0, g["model"], g["mem"], n, 0, 0))
                self.local_index.append(i)
            node_off.append(len(rows))
        self.gpus = np.array(rows, dtype=R.GPU_DTYPE) if rows else np.zeros(0, dtype=R.GPU_DTYPE)
        self.node_off = np.array(node_off, dtype=np.uint32)
        self.node_index = {n: i for i, n in enumerate(self.node_names)}

    def results(self, out: np.ndarray, request_names, profile_ids: dict) -> list:
        """OutRecs -> DeviceRequestAllocationResult dicts (types.go:795-840); failed slots -> None."""
        res = []
        for o, req in zip(out, request_names):
            if int(o["status"]) != R.ST_OK:
                res.append(None)
                continue
            g = int(o["gpu"])
            node = int(np.searchsorted(self.node_off, g, side="right") - 1)
            idx = self.local_index[g]
            if int(o["profile"]) < R.MAX_PROFILES:
                dev = canonical_mig_name(idx, profile_ids[int(o["profile"])], int(o["start"]), int(o["size"]))
            else:
                dev = canonical_name(idx)
            res.append({"request": req, "driver": DRIVER_NAME, "pool": self.node_names[node], "device": dev})
        return res


def lower_request(profile_enums: dict, *, device_class: str, profile: str | None = None, count: int = 1,
                  node: int = 0, group: int = 0, sharing_strategy: str | None = None, mem_limit=None) -> np.ndarray:
    """One DeviceRequest (+ DeviceClass + the CEL subset the driver's own specs use) -> ClaimRec.

    device_class: 'gpu.nvidia.com' | 'mig.nvidia.com' (deployments/helm/.../deviceclass-{gpu,mig}.yaml:10);
    profile: the string compared in `device.attributes['gpu.nvidia.com'].profile == '...'` (gpu-test4.yaml:23-25);
    sharing_strategy / mem_limit: opaque GpuConfig sharing (gpu-test5.yaml:24-45)."""
    c = np.zeros(1, dtype=R.CLAIM_DTYPE)
    c["node"], c["group"], c["count"] = node, group, 1
    if device_class.startswith("mig"):
        if profile not in profile_enums:
            raise KeyError(f"unknown MIG profile {profile!r}")
        c["kind"], c["profile"] = R.KIND_MIG, profile_enums[profile]
    elif sharing_strategy in ("TimeSlicing", "MPS"):
        c["kind"] = R.KIND_SHARED
        if sharing_strategy == "MPS" and mem_limit is not None:
            c["mem_limit_mib"] = sharing.megabyte_mib(mem_limit)
    else:
        c["kind"], c["count"] = R.KIND_GPU, count
    return c
