"""GPU suite: the C++ host layer (include/dra_driver.hpp — the classic controller.Driver surface with the
reference's names) driven by tests/cpp/driver_test.cpp through the C ABI."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_driver_suite(pkg):
    exe = pkg.build.build_cpp_tests()
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(pkg.api.SO_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
