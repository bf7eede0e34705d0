"""CPU suite, part 4: ResourceSlice devices <-> flat records (host-side index arithmetic)."""
import numpy as np
import pytest


def _gpu(i, mem="40Gi"):
    # GpuInfo.GetDevice, cmd/nvidia-dra-plugin/deviceinfo.go:98-142
    return {"name": f"gpu-{i}", "basic": {"attributes": {"type": {"string": "gpu"}, "index": {"int": i},
                                                            "productName": {"string": "NVIDIA A100-SXM4-40GB"}},
                                          "capacity": {"memory": {"value": mem}}}}


def _mig(parent, profile, pid, start, size):
    # MigDeviceInfo.GetDevice, deviceinfo.go:144-206 — one memorySlice<i> capacity per occupied slice
    cap = {f"memorySlice{i}": {"value": "1"} for i in range(start, start + size)}
    cap["memory"] = {"value": "5Gi"}
    return {"name": f"gpu-{parent}-mig-{pid}-{start}-{size}",
            "basic": {"attributes": {"type": {"string": "mig"}, "parentIndex": {"int": parent}, "profile": {"string": profile},
                                     "productName": {"string": "NVIDIA A100-SXM4-40GB"}}, "capacity": cap}}


def test_inventory_from_published_devices(pkg):
    C, R = pkg.codec, pkg.records
    pools = {"node-a": [_gpu(4), _gpu(5), _mig(0, "1g.5gb", 19, 0, 1), _mig(0, "2g.10gb", 14, 2, 2), _mig(1, "3g.20gb", 9, 4, 4)],
             "node-b": [_gpu(0, "80Gi")], "node-c": []}
    inv = C.Inventory(pools)
    assert list(inv.node_off) == [0, 4, 5, 5] and inv.local_index == [0, 1, 4, 5, 0]
    assert list(inv.gpus["busy"]) == [0b1101, 0xF0, 0, 0, 0]
    assert list(inv.gpus["flags"]) == [R.GPU_MIG_ENABLED, R.GPU_MIG_ENABLED, 0, 0, 0]
    assert list(inv.gpus["mem_free_mib"]) == [0, 0, 40960, 40960, 81920]
    assert list(inv.gpus["node"]) == [0, 0, 0, 0, 1]


def test_names_round_trip(pkg):
    C = pkg.codec
    assert C.canonical_name(3) == "gpu-3" and C.parse_name("gpu-3") == ("gpu", 3)
    n = C.canonical_mig_name(2, 19, 5, 1)
    assert n == "gpu-2-mig-19-5-1" and C.parse_name(n) == ("mig", 2, 19, 5, 1)
    with pytest.raises(ValueError):
        C.parse_name("imex-channel-3")


def test_lower_request_and_results(pkg, oracle):
    C, R = pkg.codec, pkg.records
    enums = {v: k for k, v in R.A100_40GB_NAMES.items()}
    ids = {R.GI_1_SLICE: 19, R.GI_2_SLICE: 14, R.GI_3_SLICE: 9, R.GI_7_SLICE: 0}
    inv = C.Inventory({"n0": [_mig(0, "1g.5gb", 19, 0, 1)], "n1": [_gpu(0), _gpu(1)]})
    reqs = [C.lower_request(enums, device_class="mig.nvidia.com", profile="1g.5gb", node=0, group=7),
            C.lower_request(enums, device_class="mig.nvidia.com", profile="3g.20gb", node=0, group=7),
            C.lower_request(enums, device_class="gpu.nvidia.com", node=1),
            C.lower_request(enums, device_class="gpu.nvidia.com", node=1, sharing_strategy="MPS", mem_limit="10Gi")]
    claims = np.concatenate(reqs)
    assert list(claims["kind"]) == [1, 1, 0, 2] and claims["mem_limit_mib"][3] == 10240
    out, _ = oracle.allocate(inv.gpus, inv.node_off, R.default_table(), claims)
    res = inv.results(out, ["mig-a", "mig-b", "gpu", "shared"], ids)
    assert res[0] == {"request": "mig-a", "driver": "gpu.nvidia.com", "pool": "n0", "device": "gpu-0-mig-19-1-1"}
    assert res[1]["device"] == "gpu-0-mig-9-4-4"
    assert res[2]["device"] == "gpu-0" and res[2]["pool"] == "n1"
    assert res[3]["device"] == "gpu-1"
    with pytest.raises(KeyError):
        C.lower_request(enums, device_class="mig.nvidia.com", profile="9g.99gb")
