"""The C-ABI boundary: the library loads, exports exactly what include/msplat_b200.h declares,
validates arguments, and the Python layer fails loudly instead of falling back."""
import os
import re

import pytest
import torch

import msplat_b200
from msplat_b200 import _lib
from conftest import ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "msplat_b200.h")).read()
    declared = set(re.findall(r"\b(msb_\w+)\s*\(", hdr))
    bound = set(_lib.exported_symbols())  # getattr on each: raises if the .so lacks one
    assert declared == bound
    assert _lib.lib().msb_version() >= 100


def test_public_api_names_match_reference():
    # /root/reference/msplat/__init__.py:11-19
    ref_names = {"project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "compute_sh", "alpha_blending",
                 "rasterization"}
    assert ref_names <= set(msplat_b200.__all__)
    assert set(msplat_b200.__all__) - ref_names == {"rasterization_sh", "rasterization_sh_views",
                                                    "sort_gaussian_views"}  # extensions
    import inspect
    sig = inspect.signature(msplat_b200.project_point)
    assert list(sig.parameters) == ["xyz", "intr", "extr", "W", "H", "nearest", "extent"]
    assert sig.parameters["nearest"].default == 0.0 and sig.parameters["extent"].default == 1.3
    assert list(inspect.signature(msplat_b200.alpha_blending).parameters) == [
        "uv", "conic", "opacity", "feature", "idx_sorted", "tile_range", "bg", "W", "H", "ndc"]
    assert list(inspect.signature(msplat_b200.rasterization).parameters)[:11] == [
        "xyz", "scale", "rotate", "opacity", "feature", "intr", "extr", "W", "H", "bg", "ndc"]


def test_no_cpu_fallback():
    """CPU tensors are rejected (reference: CHECK_CUDA, include/utils.h:9-10); nothing routes to the oracle."""
    xyz = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        msplat_b200.project_point(xyz, torch.ones(4), torch.eye(4)[:3], 64, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        msplat_b200.compute_sh(torch.zeros(4, 3, 16), torch.zeros(4, 3))
    import subprocess, sys
    out = subprocess.run([sys.executable, "-c", "import sys, msplat_b200; print('oracle' in sys.modules)"],
                         cwd=ROOT, capture_output=True, text=True)
    assert out.stdout.strip() == "False", out.stdout + out.stderr


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.msb_blend_cpad(3) == 4 and L.msb_blend_cpad(8) == 8 and L.msb_blend_cpad(33) == 48
    assert L.msb_sort_num_passes(1920, 1080) == 6      # T = 8160 -> 13 tile bits -> 45 bits
    assert L.msb_sort_num_passes(3840, 2160) == 6      # T = 32400 -> 47 bits
    assert L.msb_sort_num_passes(256, 256) == 5        # T = 256 -> 40 bits
    assert L.msb_sort_workspace_bytes(500, 1000, 64, 64) > 1000 * 12 + 500 * 24
    # negative P / null pointers are rejected before any launch
    rc = L.msb_project_point_fwd(None, None, None, 5, 64, 64, 0.0, 1.3, None, None, None)
    assert rc == -1 and b"project_point_fwd" in L.msb_last_error()
    rc = L.msb_compute_sh_fwd(None, None, None, 5, 3, 17, None, None)
    assert rc == -1


def test_binding_table_matches_header_prototypes():
    """Every prototype of include/msplat_b200.h and its ctypes binding in _lib.py agree in arity and
    in the kind of each parameter (pointer / int / float / size_t / long long): catches a signature that
    changed on one side only -- a mismatch would otherwise only show up as garbage on the GPU."""
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "msplat_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|const char\*)\s+(msb_\w+)\s*\(([^)]*)\)\s*;", hdr)
    assert len(protos) >= 27
    L = _lib.lib()

    def kind(decl):
        decl = decl.strip()
        if "*" in decl:
            return "ptr"
        base = decl.rsplit(" ", 1)[0].strip()
        return {"int": "int", "float": "float", "size_t": "size_t", "long long": "ll"}[base]

    ckind = {ctypes.c_void_p: "ptr", ctypes.c_int: "int", ctypes.c_float: "float", ctypes.c_size_t: "size_t",
             ctypes.c_longlong: "ll"}
    for name, params in protos:
        params = params.strip()
        want = [] if params in ("", "void") else [kind(p) for p in params.split(",")]
        got = [ckind[a] for a in getattr(L, name).argtypes]
        assert got == want, f"{name}: header {want} vs _lib.py {got}"
