"""CPU suite: the host side of the boundary — the CEL subset the driver's own specs use, the opaque
config precedence of device_state.go, NVML-sourced placement tables with the G == C filter, claim lowering and the
static-MIG reading of a ResourceSlice.  The spec shapes below restate demo/specs/quickstart/gpu-test{4,5,6}.yaml and
deployments/helm/k8s-dra-driver/templates/deviceclass-{gpu,mig}.yaml as Python dicts."""
import numpy as np
import pytest

from test_codec import _gpu, _mig

A = "device.attributes['gpu.nvidia.com']"
CLASSES = {"gpu.nvidia.com": [f"device.driver == 'gpu.nvidia.com' && {A}.type == 'gpu'"],          # deviceclass-gpu.yaml:10
           "mig.nvidia.com": [f"device.driver == 'gpu.nvidia.com' && {A}.type == 'mig'"]}          # deviceclass-mig.yaml:10
TEST6 = f"""{A}.productName.lowerAscii().matches('^.*a100.*$')
              &&
              ({A}.index == 0 ||
               {A}.index == 2 ||
               {A}.index == 4 ||
               {A}.index == 6)"""                                                                   # gpu-test6.yaml:23-31


def _req(name, cls, expr=None, **kw):
    r = {"name": name, "deviceClassName": cls}
    if expr:
        r["selectors"] = [{"cel": {"expression": expr}}]
    r.update(kw)
    return r


def _cfg(requests, **params):
    return {"requests": requests, "opaque": {"driver": "gpu.nvidia.com", "parameters": dict(apiVersion="gpu.nvidia.com/v1alpha1", **params)}}


def test_cel_subset(pkg):
    C, R = pkg.cel, pkg.records
    lo = C.lower(CLASSES["mig.nvidia.com"][0])
    assert (lo.kind, lo.profile, lo.program) == (R.KIND_MIG, None, [])
    lo = C.lower(f"{A}.profile == '1g.5gb'")                                                        # gpu-test4.yaml:23-25
    assert (lo.kind, lo.profile, lo.program) == (None, "1g.5gb", [])
    products = ["NVIDIA A100-SXM4-40GB", "NVIDIA H100 80GB HBM3", "NVIDIA A100 80GB PCIe", "NVIDIA B200"]
    lo = C.lower(TEST6, products)
    # the disjunction of index == k folds into ONE IN_MASK instruction: 3 instructions instead of 9
    assert lo.program == [("cmp", R.ATTR_PRODUCT, R.CMP_IN_MASK, 0b0101), ("cmp", R.ATTR_INDEX, R.CMP_IN_MASK, 0b01010101), "and"]
    assert lo.selector().dtype == R.SEL_INS_DTYPE
    lo = C.lower(f"device.capacity['gpu.nvidia.com'].memory.compareTo(quantity('40Gi')) >= 0 && "
                 f"!({A}.cudaComputeCapability.compareTo(semver('9.0.0')) < 0) && {A}.driverVersion.compareTo(semver('550.0.0')) >= 0")
    assert lo.program == [("cmp", R.ATTR_MEMORY_MIB, R.CMP_GE, 40960), ("cmp", R.ATTR_CC, R.CMP_LT, 0x0900), "not", "and",
                          ("cmp", R.ATTR_DRIVER_MAJOR, R.CMP_GE, 550), "and"]
    assert C.lower(f"3 < {A}.index").program == [("cmp", R.ATTR_INDEX, R.CMP_GT, 3)]                # literal on the left
    assert C.lower(f"{A}.productName == 'NVIDIA B200'", products).program == [("cmp", R.ATTR_PRODUCT, R.CMP_IN_MASK, 0b1000)]
    for bad in (f"{A}.uuid == 'GPU-1'", f"{A}.index == 1 || {A}.type == 'gpu'", "device.attributes['other.com'].index == 1",
                f"{A}.index ==", f"{A}.driverVersion.compareTo(semver('550.1.0')) >= 0", "device.driver == 'other'",
                " || ".join(f"({A}.index == {i} && {A}.index != {i + 40})" for i in range(5))):
        with pytest.raises(C.CelError):
            C.lower(bad, products)


def test_selector_bytecode_evaluates_like_the_expression(pkg, oracle):
    """The lowered gpu-test6 program, run by the oracle's selector interpreter over a node of 8 GPUs of two products:
    exactly the A100s with an even index pass."""
    C, R = pkg.cel, pkg.records
    products = ["NVIDIA A100-SXM4-40GB", "NVIDIA H100 80GB HBM3"]
    g, off = R.make_inventory([8], mig=False)
    attrs = np.zeros(8, dtype=R.ATTR_DTYPE)
    attrs["index"] = np.arange(8); attrs["product"] = [0, 0, 1, 0, 0, 1, 0, 0]
    sels = np.stack([C.lower(TEST6, products).selector()])
    c = np.zeros(8, dtype=R.CLAIM_DTYPE)
    c["kind"], c["count"], c["mem_limit_mib"] = R.KIND_GPU, 1, 1
    oracle.set_selectors(attrs, sels)
    try:
        out, _ = oracle.allocate(g, off, R.default_table(), c)
    finally:
        oracle.set_selectors()
    assert sorted(int(x) for x in out["gpu"][out["status"] == 0]) == [0, 4, 6]                     # index 2 is an H100
    assert (out["status"][3:] == R.ST_NO_CAPACITY).all()


def test_config_precedence(pkg):
    K = pkg.configs
    mk = lambda src, reqs, **p: dict(_cfg(reqs, **p), source=src)                                   # noqa: E731
    possible = [mk("FromClaim", ["a"], kind="GpuConfig", sharing={"strategy": "TimeSlicing"}),
                mk("FromClass", [], kind="GpuConfig", sharing={"strategy": "MPS", "mpsConfig": {"defaultPinnedDeviceMemoryLimit": "2Gi"}}),
                mk("FromClaim", ["a", "b"], kind="GpuConfig", sharing={"strategy": "MPS", "mpsConfig": {"defaultPinnedDeviceMemoryLimit": "10Gi"}}),
                {"source": "FromClaim", "requests": ["a"], "opaque": {"driver": "other.example.com", "parameters": {"kind": "X"}}},
                mk("FromClaim", ["m"], kind="MigDeviceConfig")]
    got = K.get_opaque_device_configs(possible)
    assert [c["requests"] for c in got] == [[], ["a"], ["a", "b"], ["m"]]                           # class first, list order kept, other driver skipped
    # claim beats class, later beats earlier: request a -> the LAST claim config naming it
    assert K.sharing_of(K.effective_config("a", "gpu", possible)) == ("MPS", 10240)
    assert K.sharing_of(K.effective_config("b", "gpu", possible)) == ("MPS", 10240)
    # a request no claim config names falls to the class config with an empty request list
    assert K.sharing_of(K.effective_config("zzz", "gpu", possible)) == ("MPS", 2048)
    # ... and a MIG device skips GpuConfigs without requests and ends at the MigDeviceConfig default
    assert K.effective_config("zzz", "mig", possible) == {"kind": "MigDeviceConfig", "default": True}
    # a config that NAMES the request must fit the device type (device_state.go:236-244)
    with pytest.raises(K.ConfigError):
        K.effective_config("m", "gpu", possible)
    with pytest.raises(K.ConfigError):
        K.get_opaque_device_configs([{"source": "FromNowhere", "requests": []}])
    with pytest.raises(K.ConfigError):
        K.get_opaque_device_configs([{"source": "FromClaim", "requests": [], "opaque": None}])
    # defaults are inserted at the front, GpuConfig first (device_state.go:205-221)
    assert [c["config"]["kind"] for c in K.default_configs()] == ["ImexChannelConfig", "MigDeviceConfig", "GpuConfig"]
    # the mapping loop over several results
    cfgs = K.default_configs() + got
    res = [{"request": "a", "device": "gpu-0"}, {"request": "m", "device": "gpu-1-mig-19-0-1"}, {"request": "q", "device": "gpu-2-mig-19-1-1"}]
    m = K.map_configs_to_results(res, lambda d: "mig" if "-mig-" in d else "gpu", cfgs)
    assert m == {5: [0], 6: [1], 1: [2]}
    with pytest.raises(pkg.sharing.ErrInvalidLimit):
        K.sharing_of({"kind": "GpuConfig", "sharing": {"strategy": "MPS", "mpsConfig": {"defaultPinnedDeviceMemoryLimit": "1M"}}})


def test_claims_of_the_quickstart_specs(pkg, oracle):
    C, R = pkg.codec, pkg.records
    enums = {v: k for k, v in R.A100_40GB_NAMES.items()}
    ids = {R.GI_1_SLICE: 19, R.GI_2_SLICE: 14, R.GI_3_SLICE: 9, R.GI_7_SLICE: 0}
    # gpu-test4.yaml:19-44 — four MIG requests that must share a parent
    spec4 = {"devices": {"requests": [_req("mig-1g-5gb-0", "mig.nvidia.com", f"{A}.profile == '1g.5gb'"),
                                      _req("mig-1g-5gb-1", "mig.nvidia.com", f"{A}.profile == '1g.5gb'"),
                                      _req("mig-2g-10gb", "mig.nvidia.com", f"{A}.profile == '2g.10gb'"),
                                      _req("mig-3g-20gb", "mig.nvidia.com", f"{A}.profile == '3g.20gb'")],
                         "constraints": [{"requests": [], "matchAttribute": "gpu.nvidia.com/parentUUID"}]}}
    claims, names, sels = C.lower_claim(spec4, CLASSES, enums, node=0, group=3)
    assert list(claims["kind"]) == [1] * 4 and list(claims["group"]) == [3] * 4 and sels == []
    assert list(claims["profile"]) == [R.GI_1_SLICE, R.GI_1_SLICE, R.GI_2_SLICE, R.GI_3_SLICE]
    # an EMPTY MIG-enabled GPU is in no ResourceSlice of this snapshot (nvlib.go:152, :159-171): handed in explicitly
    inv = C.Inventory({"n0": []}, extra_gpus={"n0": [0]})
    out, _ = oracle.allocate(inv.gpus, inv.node_off, R.default_table(), claims)
    assert [r["device"] for r in inv.results(out, names, ids)] == ["gpu-0-mig-19-0-1", "gpu-0-mig-19-1-1", "gpu-0-mig-14-2-2", "gpu-0-mig-9-4-4"]
    # gpu-test5.yaml:19-45 — the config list decides that both requests are SHARED, MPS with its 10Gi limit
    spec5 = {"devices": {"requests": [_req("ts-gpu", "gpu.nvidia.com"), _req("mps-gpu", "gpu.nvidia.com")],
                         "config": [_cfg(["ts-gpu"], kind="GpuConfig", sharing={"strategy": "TimeSlicing", "timeSlicingConfig": {"interval": "Long"}}),
                                    _cfg(["mps-gpu"], kind="GpuConfig", sharing={"strategy": "MPS", "mpsConfig": {
                                        "defaultActiveThreadPercentage": 50, "defaultPinnedDeviceMemoryLimit": "10Gi"}})]}}
    claims, names, _ = C.lower_claim(spec5, CLASSES, enums, node=0)
    assert list(claims["kind"]) == [R.KIND_SHARED] * 2 and list(claims["mem_limit_mib"]) == [0, 10240] and names == ["ts-gpu", "mps-gpu"]
    # gpu-test6.yaml:19-41 — CEL selector + TimeSlicing config
    spec6 = {"devices": {"requests": [_req("gpu", "gpu.nvidia.com", TEST6)],
                         "config": [_cfg(["gpu"], kind="GpuConfig", sharing={"strategy": "TimeSlicing"})]}}
    inv = C.Inventory({"n0": [_gpu(i) for i in range(4)]})
    claims, names, sels = C.lower_claim(spec6, CLASSES, enums, node=0, products=inv.products)
    assert claims["kind"][0] == R.KIND_SHARED and claims["group"][0] == 1 and len(sels) == 1       # SHARED carries its selector id in `group`
    oracle.set_selectors(inv.attrs, np.stack(sels))
    try:
        out, _ = oracle.allocate(inv.gpus, inv.node_off, R.default_table(), np.concatenate([claims] * 3))
    finally:
        oracle.set_selectors()
    assert [r["device"] for r in inv.results(out, names * 3, ids)] == ["gpu-0", "gpu-0", "gpu-0"]  # time-sliced: the same GPU again
    # a full-GPU request without sharing config keeps its count
    claims, _, _ = C.lower_claim({"devices": {"requests": [_req("g", "gpu.nvidia.com", count=2)]}}, CLASSES, enums)
    assert (claims["kind"][0], claims["count"][0]) == (R.KIND_GPU, 2)


class FakeNvml:
    """A100-40GB as NVML reports it (profile ids / placements / memory in memory-slice units, public MIG user guide) —
    SYNTHETIC, used to drive the enumeration loop without a MIG-capable GPU."""
    PROFILES = {0: (19, 4864, [(s, 1) for s in range(7)]), 1: (14, 9856, [(0, 2), (2, 2), (4, 2)]), 2: (9, 19968, [(0, 4), (4, 4)]),
                3: (5, 19968, [(0, 4)]), 4: (0, 40192, [(0, 8)]), 7: (20, 4864, [(s, 1) for s in range(7)])}

    def gpu_instance_profile_info(self, dev, i):
        if i not in self.PROFILES:
            return (3 if i < 9 else 2), None                     # NOT_SUPPORTED / INVALID_ARGUMENT: both skipped (nvlib.go:247-252)
        return 0, {"id": self.PROFILES[i][0], "memory_size_mb": self.PROFILES[i][1]}

    def gpu_instance_possible_placements(self, dev, pid):
        for i, (id_, _, pl) in self.PROFILES.items():
            if id_ == pid:
                return 0, pl
        return 3, []


def test_placement_table_from_nvml_enumeration(pkg):
    N, R = pkg.nvml_tables, pkg.records
    profs = N.enumerate_profiles(FakeNvml(), None, 40 << 30)
    # G == C (nvlib.go:283) leaves one entry per GI profile — two for the 1-slice ones, because the CI enum has TWO
    # one-slice profiles (COMPUTE_INSTANCE_PROFILE_1_SLICE and _1_SLICE_REV1, mig_profile.go:84-87): the reference lists
    # those GI profiles twice, and so does the restatement
    assert [p["name"] for p in profs] == ["1g.5gb", "1g.5gb", "2g.10gb", "3g.20gb", "4g.20gb", "7g.40gb", "1g.5gb+me", "1g.5gb+me"]
    assert sorted({p["enum"] for p in profs}) == [0, 1, 2, 3, 4, 7]
    assert all(N.CI_SLICES[p["ci_profile"]] == R.GI_COMPUTE_SLICES[p["enum"]] for p in profs)
    row = N.table_row(profs)
    want = R.default_table()[0]
    for e in (R.GI_1_SLICE, R.GI_2_SLICE, R.GI_3_SLICE, R.GI_4_SLICE, R.GI_7_SLICE, R.GI_1_SLICE_REV1):
        assert row[e] == want[e], e                                          # the synthetic model-0 table IS this enumeration
    assert N.mig_memory_gb(80 << 30, 9856) == 10 and N.mig_memory_gb(int(79.6 * 2 ** 30), 40192) == 40


def test_static_mig_inventory_only_hands_out_published_devices(pkg, oracle):
    C, R = pkg.codec, pkg.records
    ids = {R.GI_1_SLICE: 19, R.GI_2_SLICE: 14, R.GI_3_SLICE: 9, R.GI_7_SLICE: 0}
    enum_of = {v: k for k, v in ids.items()}
    pools = {"n0": [_mig(0, "1g.5gb", 19, 0, 1), _mig(0, "1g.5gb", 19, 1, 1), _mig(0, "3g.20gb", 9, 4, 4),      # mig-parted "balanced"-like
                    _mig(1, "2g.10gb", 14, 0, 2), _mig(1, "2g.10gb", 14, 2, 2), _gpu(4)],
             "n1": [_mig(0, "1g.5gb", 19, 0, 1), _mig(0, "1g.5gb", 19, 1, 1), _mig(0, "3g.20gb", 9, 4, 4)]}        # same layout as n0/gpu0: same table row
    inv = C.StaticMigInventory(pools, enum_of, allocated={"n0": ["gpu-0-mig-19-0-1"]})
    assert list(inv.gpus["model"]) == [0, 1, 0, 0] and int(inv.gpus["busy"][0]) == 1
    published = {d["name"] for ds in pools.values() for d in ds}
    c = np.zeros(8, dtype=R.CLAIM_DTYPE)
    c["kind"], c["count"], c["node"] = R.KIND_MIG, 1, 0
    c["profile"] = [R.GI_1_SLICE, R.GI_1_SLICE, R.GI_2_SLICE, R.GI_2_SLICE, R.GI_2_SLICE, R.GI_3_SLICE, R.GI_3_SLICE, R.GI_7_SLICE]
    out, _ = oracle.allocate(inv.gpus, inv.node_off, inv.table, c)
    res = inv.results(out, [f"r{i}" for i in range(8)], ids)
    got = [r["device"] if r else None for r in res]
    assert got == ["gpu-0-mig-19-1-1", None, "gpu-1-mig-14-0-2", "gpu-1-mig-14-2-2", None, "gpu-0-mig-9-4-4", None, None]
    assert all(d in published for d in got if d)
    # the classic reading (Inventory) would have carved NEW devices out of the free slices instead
    inv2 = C.Inventory(pools)
    out2, _ = oracle.allocate(inv2.gpus, inv2.node_off, R.default_table(), c[:1])
    assert inv2.results(out2, ["r0"], ids)[0]["device"] not in published
