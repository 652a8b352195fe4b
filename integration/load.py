"""Run the reference's UNMODIFIED Python wrappers (msplat/*.py) over libmsplat_b200.so.

    import integration.load as il
    msplat = il.load_reference_wrappers("/path/to/site-packages/msplat")   # the reference's package directory
    image = msplat.rasterization(...)                                       # reference Python, B200-native kernels

The wrappers import their backend as ``import msplat._C as _C`` (msplat/project_point.py:5 etc.); this loader
executes the reference's ``__init__.py`` as the top-level package ``msplat`` with ``msplat._C`` already bound
to :mod:`integration._C`, so none of the reference's compiled code is loaded.  Must run in a process that has
not imported the real ``msplat`` (the module name is the reference's own).
"""
import importlib.util
import os
import sys


def load_reference_wrappers(pkg_dir: str):
    if "msplat" in sys.modules:
        raise RuntimeError("a module named 'msplat' is already imported in this process")
    from integration import _C as shim
    init = os.path.join(pkg_dir, "__init__.py")
    spec = importlib.util.spec_from_file_location("msplat", init, submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["msplat"] = mod
    sys.modules["msplat._C"] = shim
    mod._C = shim
    try:
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop("msplat", None)
        sys.modules.pop("msplat._C", None)
        raise
    return mod
