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

from . import cel, configs
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
    """Published devices of a set of nodes, flattened.

    Two readings of a published `type == 'mig'` device exist, and this class models ONE of them:
      * this class — the classic create-on-allocate driver BASELINE.json's north_star describes: MIG devices that
        exist are OCCUPIED slices of their parent; Allocate creates new ones from the placement table, so result names
        `gpu-P-mig-<id>-<start>-<size>` may be names no ResourceSlice lists yet.  A MIG-enabled GPU with no MIG device
        cannot be seen here at all: the snapshot publishes a full GPU only when MIG is off (nvlib.go:152) and MIG
        devices only if they exist (nvlib.go:159-171), so such a GPU is in no ResourceSlice (pass it via extra_gpus).
      * StaticMigInventory (below) — the snapshot as it stands (dynamic MIG is "not yet supported with structured
        parameters", device_state.go:512-514): the published MIG devices ARE the allocatable set.
    """

    def __init__(self, pools: dict, model_of=lambda product: 0, extra_gpus: dict | None = None):
        """pools: {node name: [Device dict, ...]} (one ResourceSlice pool per node, draplugin.go:427-435).
        extra_gpus: {node name: [index, ...]} MIG-enabled GPUs that have no MIG device yet (unpublishable, see above)."""
        extra_gpus = extra_gpus or {}
        self.node_names = list(pools)
        rows, self.local_index = [], []
        self.products: list[str] = []                 # interned productName table (dra_gpu_attr.product)
        arows = []
        node_off = [0]

        def intern(name: str) -> int:
            if name not in self.products:
                self.products.append(name)
            return self.products.index(name)

        def version(v, default=(0, 0)):
            m = re.match(r"^(\d+)(?:\.(\d+))?", str(v or ""))
            return (int(m.group(1)), int(m.group(2) or 0)) if m else default
        for n, name in enumerate(self.node_names):
            gpus: dict[int, dict] = {}
            for d in pools[name]:
                t = _attr(d, "type")
                if t == "gpu":                                   # published only when MIG is off, nvlib.go:152
                    i = int(_attr(d, "index"))
                    mem = d.get("basic", {}).get("capacity", {}).get("memory", {}).get("value", "0")
                    gpus[i] = dict(mig=False, busy=0, mem=sharing.quantity_value(mem) >> 20,
                                   model=model_of(_attr(d, "productName", "")), dev=d)
                elif t == "mig":                                 # existing MIG devices of MIG-enabled parents
                    p = int(_attr(d, "parentIndex"))
                    g = gpus.setdefault(p, dict(mig=True, busy=0, mem=0, model=model_of(_attr(d, "productName", "")), dev=d))
                    g["mig"] = True
                    for cap in d.get("basic", {}).get("capacity", {}):
                        m = _SLICE.match(cap)                    # memorySlice<i>, deviceinfo.go:199-204
                        if m:
                            g["busy"] |= 1 << int(m.group(1))
            for i in extra_gpus.get(name, ()):
                gpus.setdefault(int(i), dict(mig=True, busy=0, mem=0, model=model_of("")))
            for i in sorted(gpus):
                g = gpus[i]
                rows.append((g["busy"], R.GPU_MIG_ENABLED if g["mig"] else 0, g["model"], g["mem"], n, 0, 0))
                self.local_index.append(i)
                d = g.get("dev") or {}                  # attributes a selector can test (deviceinfo.go:102-132)
                cc = version(_attr(d, "cudaComputeCapability"))
                arows.append((g["mem"], (cc[0] << 8) | cc[1], i, intern(_attr(d, "productName", "") or ""),
                              version(_attr(d, "driverVersion"))[0] & 0xFFFF))
            node_off.append(len(rows))
        self.gpus = np.array(rows, dtype=R.GPU_DTYPE) if rows else np.zeros(0, dtype=R.GPU_DTYPE)
        self.node_off = np.array(node_off, dtype=np.uint32)
        self.attrs = np.array(arows, dtype=R.ATTR_DTYPE) if arows else np.zeros(0, dtype=R.ATTR_DTYPE)
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


class StaticMigInventory(Inventory):
    """The snapshot's own model: the MIG devices a node PUBLISHES are the candidates; allocation picks
    among them and never invents a placement.  Encoding on the unchanged record layout: every distinct (base model,
    set of published placements) becomes its own placement-table row whose start masks hold exactly the published
    devices' starts, and `busy` holds the slices of the devices that are already allocated (`allocated`: device names per
    node).  First-fit over such a row can only return published devices, so every result name is in a ResourceSlice.
    At most 16 distinct layouts per context (DRA_MAX_MODELS); mig-parted style clusters have a handful."""

    def __init__(self, pools: dict, profile_enum_of_id: dict, allocated: dict | None = None, model_of=lambda product: 0):
        """profile_enum_of_id: GI profile id (the <id> of gpu-P-mig-<id>-<start>-<size>) -> NVML enum (table column)."""
        super().__init__(pools, model_of)
        allocated = allocated or {}
        layouts, self.table = {}, R.empty_table()
        for n, name in enumerate(self.node_names):
            per_gpu: dict[int, dict] = {}
            for d in pools[name]:
                if _attr(d, "type") != "mig":
                    continue
                kind, parent, pid, start, size = parse_name(d["name"])
                per_gpu.setdefault(parent, {}).setdefault((profile_enum_of_id[pid], size), set()).add(start)
            taken = {}
            for dev in allocated.get(name, ()):
                t = parse_name(dev)
                if t[0] == "mig":
                    taken[t[1]] = taken.get(t[1], 0) | (((1 << t[4]) - 1) << t[3])
                else:
                    taken[t[1]] = -1
            for gi in range(int(self.node_off[n]), int(self.node_off[n + 1])):
                idx = self.local_index[gi]
                if not (self.gpus["flags"][gi] & R.GPU_MIG_ENABLED):
                    if taken.get(idx) == -1:
                        self.gpus["flags"][gi] |= R.GPU_FULL_ALLOCATED
                    continue
                lay = per_gpu.get(idx, {})
                key = (int(self.gpus["model"][gi]), tuple(sorted((e, sz, tuple(sorted(st))) for (e, sz), st in lay.items())))
                if key not in layouts:
                    if len(layouts) >= R.MAX_MODELS:
                        raise ValueError("more than 16 distinct published MIG layouts")
                    m = len(layouts)
                    layouts[key] = m
                    for e, sz, st in key[1]:
                        self.table[m, e] = (sz, 0, R.mask_of(st))
                self.gpus["model"][gi] = layouts[key]
                self.gpus["busy"][gi] = max(0, taken.get(idx, 0))


def lower_claim(claim_spec: dict, device_classes: dict, profile_enums: dict, *, node: int = 0, group: int = 0,
                products=(), class_configs=(), selector_base: int = 0):
    """A ResourceClaim(Template) spec -> ClaimRecs, the way the Go shell's lowering would.

    claim_spec: the `spec` dict with devices.requests / constraints / config (demo/specs/quickstart/gpu-test{4,5,6}.yaml);
    device_classes: {class name: [CEL expression, ...]} (deployments/helm/k8s-dra-driver/templates/deviceclass-*.yaml:10);
    profile_enums: {profile string: NVML GI enum} of the node's GPU model; products: interned productName table.
    Returns (claims CLAIM_DTYPE[], request name per claim, [selector programs] whose 1-based ids start at selector_base+1).
    * DeviceClass + request selectors are parsed by cel.lower: `type` decides the kind, `profile ==` the profile, the
      rest becomes a selector program for the device (spec §10).
    * constraints[].matchAttribute gpu.nvidia.com/parentUUID puts the named MIG requests (all when `requests` is empty)
      into one co-location group (gpu-test4.yaml:42-44) — adjacent in the output, as spec §6 needs.
    * config (claim) + class_configs (class) go through configs.effective_config (device_state.go:226-259 precedence): a
      GpuConfig with TimeSlicing / MPS sharing turns a full-GPU request into kind SHARED with the MPS pinned-memory limit."""
    dev = claim_spec.get("devices", {})
    recs, names, sels = [], [], []
    possible = [dict(c, source="FromClass") for c in class_configs] + [dict(c, source="FromClaim") for c in dev.get("config", [])]
    grouped = set()
    for con in dev.get("constraints", []):
        if con.get("matchAttribute") == f"{cel.DOMAIN}/parentUUID":
            grouped |= set(con.get("requests") or [r["name"] for r in dev.get("requests", [])])
    for req in dev.get("requests", []):
        exprs = list(device_classes[req["deviceClassName"]]) + [s_["cel"]["expression"] for s_ in req.get("selectors", [])]
        kind, profile, program = None, None, []
        for e in exprs:
            lo = cel.lower(e, products)
            kind = lo.kind if lo.kind is not None else kind
            profile = lo.profile or profile
            if lo.program:
                program = program + lo.program + (["and"] if program else [])
        if kind is None:
            raise cel.CelError(f"request {req['name']}: the DeviceClass does not fix the device type")
        if len(program) > R.SEL_MAX_INS:
            raise cel.CelError(f"request {req['name']}: selector too long for the device")
        c = np.zeros(1, dtype=R.CLAIM_DTYPE)
        c["node"], c["count"] = node, 1
        sel_id = 0
        if program:
            prog = R.selector(*program)
            for k, p in enumerate(sels):
                if p.tobytes() == prog.tobytes():
                    sel_id = selector_base + k + 1
            if not sel_id:
                sels.append(prog)
                sel_id = selector_base + len(sels)
        count = int(req.get("count", 1)) if req.get("allocationMode", "ExactCount") == "ExactCount" else 1
        if kind == R.KIND_MIG:
            if profile is None or profile not in profile_enums:
                raise KeyError(f"request {req['name']}: unknown or missing MIG profile {profile!r}")
            c["kind"], c["profile"], c["mem_limit_mib"] = R.KIND_MIG, profile_enums[profile], sel_id
            c["group"] = group if req["name"] in grouped else 0
            copies = count
        else:
            strat, limit = configs.sharing_of(configs.effective_config(req["name"], "gpu", possible))
            if strat:
                c["kind"], c["mem_limit_mib"], c["group"] = R.KIND_SHARED, limit, sel_id
                copies = count
            else:
                c["kind"], c["count"], c["mem_limit_mib"] = R.KIND_GPU, count, sel_id
                copies = 1
        for _ in range(copies):
            recs.append(c.copy()); names.append(req["name"])
    # co-location members adjacent (stable), spec §6
    order = list(range(len(recs)))
    first_g = next((i for i in range(len(recs)) if recs[i]["group"][0] != 0), None)
    if first_g is not None:
        order = [i for i in range(len(recs)) if recs[i]["group"][0] == 0 and i < first_g] + \
                [i for i in range(len(recs)) if recs[i]["group"][0] != 0] + \
                [i for i in range(len(recs)) if recs[i]["group"][0] == 0 and i > first_g]
    claims = np.concatenate([recs[i] for i in order]) if recs else np.zeros(0, dtype=R.CLAIM_DTYPE)
    return claims, [names[i] for i in order], sels
