"""Opaque device configs: which config applies to which allocation result.

Restates, on plain dicts, the two pieces of cmd/nvidia-dra-plugin/device_state.go that decide it:
  * GetOpaqueDeviceConfigs (device_state.go:457-510): class configs before claim configs (so claim beats class when the
    list is walked backwards), list order kept inside each source, other drivers' configs skipped, non-opaque configs and
    unknown sources are errors;
  * the mapping loop of prepareDevices (device_state.go:205-259): three defaults (GpuConfig, MigDeviceConfig,
    ImexChannelConfig, each with an empty request list) inserted at the FRONT in that order; for every result the configs
    are walked from highest to lowest precedence; a config that names the result's request wins (a type mismatch is an
    error), a config with no requests wins if its type fits the device type (a mismatch is skipped).
On this path it answers one question before allocation: is a request for a full GPU to be handed out SHARED
(TimeSlicing / MPS: spec §7, BASELINE configs[3]) and with which pinned-memory limit (sharing.go:190-237).
Host-side only.
"""
from __future__ import annotations

from . import sharing

DRIVER_NAME = "gpu.nvidia.com"
GPU_CONFIG, MIG_CONFIG, IMEX_CONFIG = "GpuConfig", "MigDeviceConfig", "ImexChannelConfig"
_TYPE_OF_KIND = {GPU_CONFIG: "gpu", MIG_CONFIG: "mig", IMEX_CONFIG: "imex-channel"}     # types.go: GpuDeviceType / MigDeviceType / ImexChannelType


class ConfigError(ValueError):
    pass


def get_opaque_device_configs(possible_configs, driver_name: str = DRIVER_NAME):
    """possible_configs: list of DeviceAllocationConfiguration dicts {source, requests, opaque: {driver, parameters}}.
    Returns [{requests, config}] from lowest to highest precedence (device_state.go:457-510)."""
    class_cfgs, claim_cfgs = [], []
    for c in possible_configs:
        src = c.get("source")
        if src == "FromClass":
            class_cfgs.append(c)
        elif src == "FromClaim":
            claim_cfgs.append(c)
        else:
            raise ConfigError(f"invalid config source: {src}")
    out = []
    for c in class_cfgs + claim_cfgs:
        op = c.get("opaque")
        if op is None:
            raise ConfigError("only opaque parameters are supported by this driver")
        if op.get("driver") != driver_name:
            continue                                   # another driver's config for the same request: not an error
        params = op.get("parameters")
        if not isinstance(params, dict) or params.get("kind") not in _TYPE_OF_KIND:
            raise ConfigError(f"error decoding config parameters: {params!r}")
        out.append({"requests": list(c.get("requests") or []), "config": params})
    return out


def default_configs():
    """device_state.go:205-221: each default is inserted at index 0, GpuConfig first — so the list starts
    [ImexChannelConfig, MigDeviceConfig, GpuConfig]."""
    cfgs = []
    for kind in (GPU_CONFIG, MIG_CONFIG, IMEX_CONFIG):
        cfgs.insert(0, {"requests": [], "config": {"kind": kind, "default": True}})
    return cfgs


def map_configs_to_results(results, device_type_of, configs):
    """results: [{request, device, ...}]; device_type_of(device name) -> 'gpu' | 'mig' | 'imex-channel'.
    Returns {index into configs: [indices into results]} (device_state.go:223-259)."""
    out = {}
    for ri, res in enumerate(results):
        dtype = device_type_of(res["device"])
        if dtype is None:
            raise ConfigError(f"requested device is not allocatable: {res['device']}")
        for ci in range(len(configs) - 1, -1, -1):
            c = configs[ci]
            ctype = _TYPE_OF_KIND[c["config"]["kind"]]
            if res["request"] in c["requests"]:
                if ctype != dtype:
                    raise ConfigError(f"cannot apply {c['config']['kind']} to request: {res['request']}")
                out.setdefault(ci, []).append(ri)
                break
            if not c["requests"]:
                if ctype != dtype:
                    continue
                out.setdefault(ci, []).append(ri)
                break
    return out


def effective_config(request: str, device_type: str, possible_configs, driver_name: str = DRIVER_NAME):
    """The config that WILL govern a request's devices, decided before allocation from the request name and the device
    type its DeviceClass selects — the same walk as map_configs_to_results for one (request, type)."""
    cfgs = default_configs() + get_opaque_device_configs(possible_configs, driver_name)
    got = map_configs_to_results([{"request": request, "device": "x"}], lambda _d: device_type, cfgs)
    (ci, _), = got.items()
    return cfgs[ci]["config"]


def sharing_of(config: dict):
    """(strategy, mem_limit_mib) of a GpuConfig / MigDeviceConfig: strategy None | 'TimeSlicing' | 'MPS';
    mem_limit_mib from MpsConfig.defaultPinnedDeviceMemoryLimit by limit.Megabyte's integer rule (sharing.go:234-237);
    a limit that normalises to 0 MiB is the reference's ErrInvalidLimit."""
    sh = config.get("sharing") or {}
    strat = sh.get("strategy")
    if strat not in (None, "TimeSlicing", "MPS"):
        raise ConfigError(f"unknown GPU sharing strategy: {strat}")
    if strat != "MPS":
        return strat, 0
    lim = (sh.get("mpsConfig") or {}).get("defaultPinnedDeviceMemoryLimit")
    if lim is None:
        return strat, 0
    return strat, sharing.megabyte_mib(lim)
