"""The one piece of this path that real hardware can pin: the GPU-instance profile / placement enumeration
(cmd/nvidia-dra-plugin/nvlib.go:244-295).  tests/golden/b200_gi_profiles.json holds what NVML answered on the B200 box
(recorded by tests/golden/make_b200_gi_profiles.py) next to `nvidia-smi mig -lgip / -lgipp` of the same GPU.  Here:
the recorded answers replayed through nvml_tables.enumerate_profiles, the names (go-nvlib mig_profile.go arithmetic), ids
and placements compared with nvidia-smi's own listing, and the allocator run on the resulting dra_profile_tbl row."""
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def doc():
    return json.load(open(os.path.join(GOLDEN, "b200_gi_profiles.json")))


class _Replay:
    """The two NVML calls, answered from the recording."""

    def __init__(self, doc):
        self.by_profile = {e["profile"]: e for e in doc["nvml"]}
        self.by_id = {e["info"]["id"]: e for e in doc["nvml"] if e["info_ret"] == 0}

    def gpu_instance_profile_info(self, _dev, i):
        e = self.by_profile[i]
        return e["info_ret"], e.get("info", {})

    def gpu_instance_possible_placements(self, _dev, pid):
        e = self.by_id[pid]
        return e["placements_ret"], [tuple(p) for p in e["placements"]]


def _smi_profiles(text):
    """{name: id} from `nvidia-smi mig -lgip`."""
    return {m.group(1): int(m.group(2)) for m in re.finditer(r"MIG\s+(\S+)\s+(\d+)\s+\d+/\d+", text)}


def _smi_placements(text):
    """{id: (starts, size)} from `nvidia-smi mig -lgipp`."""
    out = {}
    for m in re.finditer(r"Profile ID\s+(\d+)\s+Placements?\s*:\s*\{([0-9,]*)\}:(\d+)", text):
        out[int(m.group(1))] = (tuple(int(x) for x in m.group(2).split(",")), int(m.group(3)))
    return out


def test_enumeration_replays_the_recording(pkg, doc):
    N = pkg.nvml_tables
    profs = N.enumerate_profiles(_Replay(doc), None, doc["memory_total_bytes"])
    norm = lambda ps: [(p["enum"], p["id"], p["name"], [tuple(x) for x in p["placements"]], p["ci_profile"]) for p in ps]
    assert norm(profs) == norm(doc["enumerated"])
    row = N.table_row(profs)
    assert [[int(e["size"]), int(e["start_mask"])] for e in row] == doc["row"]
    # profiles NVML does not offer on this part are skipped, not errors (nvlib.go:247-252)
    assert {e["profile"] for e in doc["nvml"] if e["info_ret"] == N.NVML_ERROR_NOT_SUPPORTED} == {5, 6, 8}


def test_names_ids_and_placements_agree_with_nvidia_smi(pkg, doc):
    """External check: nvidia-smi prints the same names (so getMigMemorySizeGB's rounding is right: 20992 MB of
    183359 MiB -> '23gb', 45312 -> '45gb', 91136 -> '90gb', 182784 -> '180gb'), ids and {starts}:size."""
    smi, smi_pl = _smi_profiles(doc["nvidia_smi_lgip"]), _smi_placements(doc["nvidia_smi_lgipp"])
    assert len(smi) == 7 and len(smi_pl) == 7
    ours = {p["name"]: p["id"] for p in doc["enumerated"]}
    assert ours == smi
    for p in doc["enumerated"]:
        starts, size = smi_pl[p["id"]]
        assert tuple(st for st, _ in p["placements"]) == starts and {sz for _, sz in p["placements"]} == {size}
    # the G == C loop lists the 1-compute-slice profiles twice (the CI enum has two 1-slice profiles, nvlib.go:276-292)
    names = [p["name"] for p in doc["enumerated"]]
    assert names.count("1g.23gb") == 2 and names.count("1g.45gb") == 2 and names.count("3g.90gb") == 1


def test_b200_row_has_the_geometry_of_the_synthetic_model0_row(pkg, doc):
    """The spec's synthetic model-0 table (A100-40GB geometry, written from memory in round 1) against real hardware:
    the B200 reports the same {size, starts} for every profile both have."""
    R = pkg.records
    for enum, (size, starts) in R.a100_40gb_rows().items():
        assert doc["row"][enum] == [size, R.mask_of(starts)], enum


def _b200_workload(pkg, doc, n_claim, n_node, seed):
    """Mixed MIG claims over every profile the B200 offers, on nodes of 8 B200s (model 15 = the recorded row)."""
    R = pkg.records
    t = R.default_table()
    for p, (size, mask) in enumerate(doc["row"]):
        t[15, p] = (size, 0, mask)
    g, off = R.make_inventory([8] * n_node, model=15, mem_free_mib=183359)
    r = pkg.synth.splitmix64(0xB200 + seed, 3 * n_claim)
    offered = np.array([p for p, (size, _) in enumerate(doc["row"]) if size], dtype=np.uint8)
    c = np.zeros(n_claim, dtype=R.CLAIM_DTYPE)
    c["kind"] = R.KIND_MIG
    c["profile"] = offered[(r[0::3] % np.uint64(len(offered))).astype(np.int64)]
    c["count"] = 1
    c["node"] = (r[1::3] % np.uint64(n_node)).astype(np.uint32)
    pre = (r[2::3][: len(g)] % np.uint64(4) == 0)                    # a quarter of the GPUs start with slices 2-3 taken
    g["busy"][pre] = 0b1100
    return pkg.synth.Workload(f"b200_{seed}", g, off, t, c).finish()


def test_known_answers_on_an_empty_b200(pkg, oracle, doc):
    R = pkg.records
    t = R.default_table()
    for p, (size, mask) in enumerate(doc["row"]):
        t[15, p] = (size, 0, mask)
    g, off = R.make_inventory([1], model=15)

    def run(profiles):
        c = np.zeros(len(profiles), dtype=R.CLAIM_DTYPE)
        c["kind"], c["count"], c["profile"] = R.KIND_MIG, 1, profiles
        out, _ = oracle.allocate(g, off, t, c)
        return [(int(o["start"]), int(o["status"])) for o in out]

    # nvidia-smi: 1g.23gb 7/7 instances, 1g.45gb 4/4, 2g.45gb 3/3, 3g.90gb 2/2, 4g.90gb 1/1, 7g.180gb 1/1
    assert run([R.GI_1_SLICE] * 8) == [(s, 0) for s in range(7)] + [(0, R.ST_NO_CAPACITY)]
    assert run([R.GI_1_SLICE_REV2] * 5) == [(0, 0), (2, 0), (4, 0), (6, 0), (0, R.ST_NO_CAPACITY)]
    assert run([R.GI_2_SLICE] * 4) == [(0, 0), (2, 0), (4, 0), (0, R.ST_NO_CAPACITY)]
    assert run([R.GI_3_SLICE] * 3) == [(0, 0), (4, 0), (0, R.ST_NO_CAPACITY)]
    assert run([R.GI_4_SLICE, R.GI_4_SLICE]) == [(0, 0), (0, R.ST_NO_CAPACITY)]
    assert run([R.GI_7_SLICE, R.GI_1_SLICE]) == [(0, 0), (0, R.ST_NO_CAPACITY)]
    assert run([R.GI_4_SLICE, R.GI_3_SLICE, R.GI_1_SLICE]) == [(0, 0), (4, 0), (0, R.ST_NO_CAPACITY)]
    assert run([R.GI_8_SLICE])[0][1] != 0                            # not offered by this part (NVML: NOT_SUPPORTED)


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_matches_naive_on_the_b200_table(pkg, oracle, doc, seed):
    from oracle import naive
    w = _b200_workload(pkg, doc, 600, 5, seed)
    out, after = oracle.allocate(w.gpus, w.node_off, w.table, w.claims)
    tbl = [[(int(e["size"]), int(e["start_mask"])) for e in row] for row in w.table]
    dicts = lambda a: [{k: int(x[k]) for k in a.dtype.names} for x in a]
    nout, nafter = naive.allocate(dicts(w.gpus), [int(x) for x in w.node_off], tbl, dicts(w.claims))
    assert [(int(o["gpu"]), int(o["start"]), int(o["size"]), int(o["profile"]), int(o["status"])) for o in out] == [tuple(x) for x in nout]
    assert int((out["status"] == 0).sum()) > 50 and int((out["status"] != 0).sum()) > 50


@pytest.mark.gpu
@pytest.mark.parametrize("n_claim,n_node,cfg", [(4000, 40, 0), (30000, 300, 0), (4000, 40, "nofused")])
def test_cuda_matches_oracle_on_the_b200_table(pkg, oracle, doc, n_claim, n_node, cfg):
    """The CUDA path on the table row a B200 deployment would load (every profile the part offers, fragmented GPUs)."""
    w = _b200_workload(pkg, doc, n_claim, n_node, 7)
    ref, ref_inv = oracle.allocate(w.gpus, w.node_off, w.table, w.claims)
    with pkg.api.Context(device=0, flags=pkg.api.CFG_NO_FUSED if cfg == "nofused" else 0) as ctx:
        ctx.set_table(w.table); ctx.set_inventory(w.gpus, w.node_off)
        out = ctx.allocate(w.claims, w.out_off, w.n_out)
        assert out.tobytes() == ref.tobytes()
        assert ctx.get_inventory().tobytes() == ref_inv.tobytes()


@pytest.mark.gpu
def test_live_nvml_matches_the_recording(pkg, doc):
    """On the GPU box: NVML answers today what the fixture recorded (same part, same driver family)."""
    N = pkg.nvml_tables
    try:
        row, profs = N.table_from_nvml(0)
    except (OSError, RuntimeError) as e:
        pytest.skip(f"NVML unavailable: {e}")
    if not profs:
        pytest.skip("this GPU offers no MIG profiles")
    assert [[int(e["size"]), int(e["start_mask"])] for e in row] == doc["row"]
    assert [(p["name"], p["id"]) for p in profs] == [(p["name"], p["id"]) for p in doc["enumerated"]]
