import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def hostsim():
    """Test-only host build of the __host__ __device__ device math (tests/hostsim)."""
    d = os.path.join(ROOT, "tests", "hostsim")
    so, src = os.path.join(d, "libhostsim.so"), os.path.join(d, "hostsim.cpp")
    deps = [src] + [os.path.join(ROOT, "vmc_pde_b200", "csrc", f) for f in ("flow_core.cuh", "flow_meta.hpp", "dc_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.check_call(["g++", "-O2", "-march=native", "-shared", "-fPIC", "-std=c++17", "-o", so, src])
    return ctypes.CDLL(so)
