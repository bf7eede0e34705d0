"""k8s-dra-driver_b200 — B200-native allocation hot path of a Kubernetes GPU DRA driver.

Import with ``importlib.import_module("k8s-dra-driver_b200")`` (the directory name is fixed by the
build contract and is not a Python identifier).  The CUDA library is loaded on first use of
``api.Context`` / ``driver.Driver``; there is no CPU fallback — a missing libdra_alloc.so or GPU raises.
"""
from . import api, build, cel, codec, configs, nvml_tables, records, shard, sharing, synth  # noqa: F401
