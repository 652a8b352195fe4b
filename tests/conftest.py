"""Shared pytest plumbing.

Markers: ``gpu`` = needs a CUDA device (the parity tests proper, run on the B200 box);
everything else runs on a CPU-only box in a few minutes.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Every gradient comparison of the GPU parity tests records how many multiples of its measured noise
    floor it needed (tests/test_gpu_parity.py::grad_close): kept as evidence for the chosen K_NOISE."""
    mod = sys.modules.get("test_gpu_parity")
    rep = getattr(mod, "REPORT", None) if mod else None
    if rep:
        import json
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "tolerance_report.json"), "w") as f:
            json.dump({"rel": mod.REL, "k_noise": mod.K_NOISE, "ulp_floor": mod.ULP_FLOOR, "checks": rep}, f, indent=1)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def host_check():
    """tools/host_check.cu compiled for the host: the product's math headers on the CPU."""
    out_dir = os.path.join(ROOT, "tools", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhost_check.so")
    src = os.path.join(ROOT, "tools", "host_check.cu")
    csrc = os.path.join(ROOT, "msplat_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        nvcc = "/usr/local/cuda/bin/nvcc"
        subprocess.check_call([nvcc, "-std=c++17", "-O1", "-arch=sm_100a", "-Xcompiler", "-fPIC,-ffp-contract=off",
                               "-shared", "-I", csrc, src, "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def ref_msplat():
    """The UNMODIFIED reference build (baseline/_ref), if it travelled with the snapshot."""
    p = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(p, "msplat")):
        pytest.skip("baseline/_ref not present (reference CUDA build not installed)")
    if p not in sys.path:
        sys.path.insert(0, p)
    try:
        import msplat  # noqa
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference build not importable: {e}")
    return msplat


def fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)
