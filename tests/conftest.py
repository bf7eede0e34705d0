import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("k8s-dra-driver_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, f"alloc_{name}.npz"))
    out_off = d["out_off"]
    return dict(gpus=d["gpus"], node_off=d["node_off"], table=d["table"], claims=d["claims"],
                out_off=(out_off if len(out_off) else None), out=d["out"], gpus_after=d["gpus_after"])


GOLDEN_CASES = ["gpu_test4", "cfg1", "cfg2_small", "cfg4_small", "cfg5_small", "mixed0", "mixed1", "mixed2",
                "mixed3", "mixed11_valid", "homog_w8", "homog_w32", "homog_16slice", "b200_table"]


@pytest.fixture(scope="session", params=["auto", "copy-engine", "bucket+pack"])
def ctx(pkg, request):
    """One CUDA context per kernel path for the GPU tests (fails loudly when the library or the device is
    missing): "auto" takes the single-launch fused kernel whenever the batch is small enough, with direct host
    I/O (the kernel reads / writes the pinned host buffers itself) where that applies; "copy-engine" is the same
    without direct host I/O (H2D / D2H copies around the kernel); "bucket+pack" forces the stable counting sort +
    pack chain."""
    flags = {"auto": 0, "copy-engine": pkg.api.CFG_NO_DIRECT, "bucket+pack": pkg.api.CFG_NO_FUSED}[request.param]
    c = pkg.api.Context(device=0, flags=flags)
    c.path = request.param
    yield c
    c.close()
