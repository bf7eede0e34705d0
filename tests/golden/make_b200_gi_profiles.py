#!/usr/bin/env python
"""Records what NVML answers on the GPU box for the two calls of the reference's enumeration loop
(cmd/nvidia-dra-plugin/nvlib.go:244-295), next to `nvidia-smi mig -lgip / -lgipp` of the same GPU:

    gpurun -- 'python tests/golden/make_b200_gi_profiles.py > gpurun_out/b200_gi_profiles.json'
    cp gpurun_out/b200_gi_profiles.json tests/golden/

tests/test_nvml_golden.py replays the recorded answers through nvml_tables.enumerate_profiles (CPU), checks names,
ids and placements against the nvidia-smi text, and runs the allocator on the resulting table row."""
import importlib, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("k8s-dra-driver_b200")
N = pkg.nvml_tables


def sh(*cmd):
    try:
        return subprocess.run(cmd, capture_output=True, text=True, timeout=60).stdout
    except Exception as e:                                                       # noqa: BLE001
        return f"<{e}>"


n = N.Nvml()
dev = n.device(0)
raw = []
for i in range(N.GPU_INSTANCE_PROFILE_COUNT):
    rc, info = n.gpu_instance_profile_info(dev, i)
    ent = {"profile": i, "info_ret": rc}
    if rc == N.NVML_SUCCESS:
        prc, pl = n.gpu_instance_possible_placements(dev, info["id"])
        ent.update(info=info, placements_ret=prc, placements=pl)
    raw.append(ent)
total = n.memory_total(dev)
profs = N.enumerate_profiles(n, dev, total)
row = N.table_row(profs)
doc = {
    "source": "tests/golden/make_b200_gi_profiles.py on the gpurun B200 box",
    "gpu": sh("nvidia-smi", "--query-gpu=name,driver_version,memory.total,mig.mode.current", "--format=csv,noheader").strip(),
    "memory_total_bytes": total,
    "mig_mode": list(n.mig_mode(dev)),
    "nvml": raw,
    "enumerated": profs,
    "row": [[int(e["size"]), int(e["start_mask"])] for e in row],
    "nvidia_smi_lgip": sh("nvidia-smi", "mig", "-i", "0", "-lgip"),
    "nvidia_smi_lgipp": sh("nvidia-smi", "mig", "-i", "0", "-lgipp"),
}
n.close()
print(json.dumps(doc, indent=1))
