"""INTEGRATION.md is code a maintainer copies: every `msb_*` call and every ctypes `argtypes` list in its
code blocks must agree with the prototypes of include/msplat_b200.h (arity, and parameter kinds for
argtypes), and the shipped binding integration/_C.py must export the 12 names of the reference's pybind11
module (/root/reference/msplat/src/ext.cpp:14-25) with the reference's parameter lists."""
import inspect
import os
import re

from conftest import ROOT

EXT_CPP_NAMES = {  # ext.cpp:14-25 and the prototypes in /root/reference/msplat/include/*.h
    "project_point_forward": ["xyz", "intr", "extr", "W", "H", "nearest", "extent"],
    "project_point_backward": ["xyz", "intr", "extr", "W", "H", "uv", "depth", "dL_duv", "dL_ddepth"],
    "compute_cov3d_forward": ["scales", "uquats", "visible"],
    "compute_cov3d_backward": ["scales", "uquats", "visible", "dL_dcov3Ds"],
    "ewa_project_forward": ["xyz", "cov3d", "intr", "extr", "uv", "W", "H", "visible"],
    "ewa_project_backward": ["xyz", "cov3d", "intr", "extr", "radius", "dL_dconic"],
    "compute_gaussian_key": ["uv", "depth", "W", "H", "radius", "tiles"],
    "compute_tile_gaussian_range": ["W", "H", "tiles", "key_sorted"],
    "compute_sh_forward": ["shs", "view_dirs", "visible"],
    "compute_sh_backward": ["shs", "view_dirs", "visible", "dL_dvalue"],
    "alpha_blending_forward": ["uv", "conic", "opacity", "feature", "idx_sorted", "tile_range", "bg", "W", "H"],
    "alpha_blending_backward": ["uv", "conic", "opacity", "feature", "idx_sorted", "tile_range", "bg", "W", "H",
                                "final_T", "ncontrib", "dL_drendered"],
}


def header_prototypes():
    hdr = open(os.path.join(ROOT, "include", "msplat_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for name, params in re.findall(r"\b(?:int|size_t|const char\*)\s+(msb_\w+)\s*\(([^)]*)\)\s*;", hdr):
        params = params.strip()
        kinds = []
        if params not in ("", "void"):
            for p in params.split(","):
                p = p.strip()
                kinds.append("V" if "*" in p else {"int": "I", "float": "F", "size_t": "SZ", "long long": "LL"}[
                    p.rsplit(" ", 1)[0].strip()])
        out[name] = kinds
    return out


def split_args(s):
    """top-level comma split of a call's argument text"""
    args, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            args.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        args.append(cur.strip())
    return args


def call_args(code, start):
    """argument text of the call whose '(' is at code[start]"""
    depth = 0
    for i in range(start, len(code)):
        if code[i] == "(":
            depth += 1
        elif code[i] == ")":
            depth -= 1
            if depth == 0:
                return code[start + 1:i]
    raise AssertionError("unbalanced call")


def test_integration_md_calls_match_the_header():
    protos = header_prototypes()
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    assert blocks
    ncalls = ntypes = 0
    for code in blocks:
        for m in re.finditer(r"\bL\.(msb_\w+)\s*\(", code):
            name = m.group(1)
            assert name in protos, f"INTEGRATION.md calls {name}, which the header does not declare"
            args = split_args(call_args(code, m.end() - 1))
            assert len(args) == len(protos[name]), \
                f"INTEGRATION.md: {name} called with {len(args)} arguments, the header declares {len(protos[name])}"
            ncalls += 1
        for m in re.finditer(r"L\.(msb_\w+)\.argtypes\s*=\s*(?:SZ,\s*)?\[([^\]]*)\]", code):
            name, lst = m.group(1), [a.strip() for a in m.group(2).split(",") if a.strip()]
            assert lst == protos[name], f"INTEGRATION.md: argtypes of {name} {lst} vs header {protos[name]}"
            ntypes += 1
    assert ncalls >= 6 and ntypes >= 4


def test_shipped_binding_mirrors_ext_cpp():
    from integration import _C
    assert set(_C.__all__) == set(EXT_CPP_NAMES)
    for name, params in EXT_CPP_NAMES.items():
        got = list(inspect.signature(getattr(_C, name)).parameters)
        assert got == params, f"integration._C.{name}{got} vs the reference's {params}"
    # every C-ABI symbol the binding uses exists in the header
    protos = header_prototypes()
    src = open(os.path.join(ROOT, "integration", "_C.py")).read()
    used = set(re.findall(r"\.(msb_\w+)", src))
    assert used and used <= set(protos), used - set(protos)


def test_loader_runs_reference_wrappers_on_cpu_box():
    """The reference's unmodified wrappers import over the binding (no compiled reference code needed); a
    call without a GPU fails loudly in OUR layer."""
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "baseline", "_ref", "msplat")
    if not os.path.isdir(pkg):
        import pytest
        pytest.skip("baseline/_ref not present")
    code = (f"import sys; sys.path.insert(0, {ROOT!r}); import integration.load as il, torch; "
            f"m = il.load_reference_wrappers({pkg!r}); import msplat; assert msplat._C.__name__ == 'integration._C'; "
            "assert not any('msplat/_C' in l for l in open('/proc/self/maps')); "
            "\ntry:\n    m.project_point(torch.zeros(4, 3), torch.ones(4), torch.eye(4)[:3], 64, 64)\n"
            "except RuntimeError as e:\n    print('RAISED', e)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "RAISED" in r.stdout and "CUDA" in r.stdout, r.stdout + r.stderr
