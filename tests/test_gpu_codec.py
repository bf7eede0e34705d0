"""GPU suite: the codec round trip on the device (VERDICT r01 missing #6) — published Device dicts -> flat records ->
dra_allocate_batch on the GPU -> DeviceRequestAllocationResult names (cmd/nvidia-dra-plugin/deviceinfo.go:74-80,
98-206; vendor/k8s.io/api/resource/v1beta1/types.go:795-840), with claims lowered from the quickstart spec shapes
(CEL subset + config precedence) and a selector program compiled from the gpu-test6 expression."""
import numpy as np
import pytest

from test_codec import _gpu, _mig
from test_host_lowering import A, CLASSES, TEST6, _cfg, _req

pytestmark = pytest.mark.gpu


def test_resource_slice_to_results_on_the_device(pkg, ctx, oracle):
    C, R = pkg.codec, pkg.records
    enums = {v: k for k, v in R.A100_40GB_NAMES.items()}
    ids = {R.GI_1_SLICE: 19, R.GI_2_SLICE: 14, R.GI_3_SLICE: 9, R.GI_7_SLICE: 0}
    pools = {"node-a": [_mig(0, "1g.5gb", 19, 0, 1), _mig(1, "3g.20gb", 9, 4, 4), _gpu(4), _gpu(5), _gpu(6)],
             "node-b": [_gpu(i) for i in range(8)]}
    for d in pools["node-b"][2::3]:
        d["basic"]["attributes"]["productName"] = {"string": "NVIDIA H100 80GB HBM3"}
    inv = C.Inventory(pools, extra_gpus={"node-a": [2]})
    spec4 = {"devices": {"requests": [_req("mig-1g-5gb-0", "mig.nvidia.com", f"{A}.profile == '1g.5gb'"),
                                      _req("mig-2g-10gb", "mig.nvidia.com", f"{A}.profile == '2g.10gb'"),
                                      _req("mig-3g-20gb", "mig.nvidia.com", f"{A}.profile == '3g.20gb'")],
                         "constraints": [{"requests": [], "matchAttribute": "gpu.nvidia.com/parentUUID"}]}}
    spec5 = {"devices": {"requests": [_req("ts-gpu", "gpu.nvidia.com"), _req("mps-gpu", "gpu.nvidia.com")],
                         "config": [_cfg(["ts-gpu"], kind="GpuConfig", sharing={"strategy": "TimeSlicing"}),
                                    _cfg(["mps-gpu"], kind="GpuConfig", sharing={"strategy": "MPS", "mpsConfig": {"defaultPinnedDeviceMemoryLimit": "10Gi"}})]}}
    spec6 = {"devices": {"requests": [_req("gpu", "gpu.nvidia.com", TEST6)]}}
    parts, names, sels = [], [], []
    for spec, node, grp in ((spec4, 0, 1), (spec5, 0, 0), (spec6, 1, 0), (spec6, 1, 0), (spec6, 1, 0), (spec6, 1, 0)):
        c, n, s_ = C.lower_claim(spec, CLASSES, enums, node=node, group=grp, products=inv.products, selector_base=0)
        for k, prog in enumerate(s_):                                  # one shared selector table: ids are 1-based into it
            if not any(prog.tobytes() == q.tobytes() for q in sels):
                sels.append(prog)
        parts.append(c); names += n
    claims = np.concatenate(parts)
    ctx.set_table(R.default_table()); ctx.set_inventory(inv.gpus, inv.node_off)
    ctx.set_gpu_attrs(inv.attrs); ctx.set_selectors(np.stack(sels))
    oracle.set_selectors(inv.attrs, np.stack(sels))
    try:
        out = ctx.allocate(claims)
        ref, _ = oracle.allocate(inv.gpus, inv.node_off, R.default_table(), claims)
    finally:
        oracle.set_selectors(); ctx.set_selectors(None); ctx.set_gpu_attrs(None)
    assert out.tobytes() == ref.tobytes()
    res = inv.results(out, names, ids)
    got = [(r["pool"], r["request"], r["device"]) if r else None for r in res]
    # GPU 0 holds one 1g device on slice 0: the run still fits there (1g@1, 2g@2-3, 3g@4-7) — lowest parent wins (spec §6)
    assert got[:3] == [("node-a", "mig-1g-5gb-0", "gpu-0-mig-19-1-1"), ("node-a", "mig-2g-10gb", "gpu-0-mig-14-2-2"),
                       ("node-a", "mig-3g-20gb", "gpu-0-mig-9-4-4")]
    assert got[3:5] == [("node-a", "ts-gpu", "gpu-4"), ("node-a", "mps-gpu", "gpu-4")]
    # gpu-test6: A100s with an even index; index 2 is an H100 (every 3rd from 2: 2, 5), so 0, 4, 6 — the 4th request fails
    assert [g[2] if g else None for g in got[5:]] == ["gpu-0", "gpu-4", "gpu-6", None]
    # results -> configs (a10): the MPS config governs exactly the mps-gpu result
    K = pkg.configs
    cfgs = K.default_configs() + K.get_opaque_device_configs([dict(c, source="FromClaim") for c in spec5["devices"]["config"]])
    m = K.map_configs_to_results([r for r in res[3:5]], lambda d: "mig" if "-mig-" in d else "gpu", cfgs)
    assert m == {3: [0], 4: [1]} and K.sharing_of(cfgs[4]["config"]) == ("MPS", 10240)


def test_nvml_placement_table_on_this_gpu(pkg):
    """The real NVML enumeration (nvlib.go:244-295) on the box's GPU.  On a GPU with MIG support the rows are loaded as
    model 15 and cross-checked for consistency; without it (or without permission) the enumeration yields nothing — both
    are reported, neither is a failure of the loop itself."""
    N, R = pkg.nvml_tables, pkg.records
    try:
        row, profs = N.table_from_nvml(0)
    except (OSError, RuntimeError) as e:
        pytest.skip(f"NVML unavailable: {e}")
    print("NVML GI profiles:", [(p["enum"], p["id"], p["name"], p["placements"]) for p in profs])
    for p in profs:
        assert p["placements"] and all(st + sz <= 16 for st, sz in p["placements"])
        e = row[p["enum"]]
        assert int(e["size"]) == p["placements"][0][1] and int(e["start_mask"]) == R.mask_of(st for st, _ in p["placements"])
    if profs:
        t = R.default_table(); t[15] = row
        with pkg.api.Context(device=0) as c:
            c.set_table(t)                                          # the library accepts the row (start + size <= 16 etc.)
