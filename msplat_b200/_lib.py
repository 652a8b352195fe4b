"""ctypes binding of libmsplat_b200.so (the C ABI declared in include/msplat_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a
tensor is not on a CUDA device the call raises.  PyTorch is used for device memory, streams
and autograd plumbing only; every kernel is ours.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_float, c_int, c_int32, c_longlong, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsplat_b200.so")

_lib = None
_lock = threading.Lock()

# Per-process launch counter: bench.py reports it as "gpu_launches" (updated under _count_lock).
_launches = 0
_count_lock = threading.Lock()

# kernels launched by each C-ABI entry point (see csrc/*.cu); a tuple means (fixed, per-pass)
_KERNELS_PER_CALL = {}


def _declare(lib):
    P, F, I, V, SZ, LL = c_void_p, c_float, c_int, c_void_p, c_size_t, c_longlong
    sig = {
        "msb_version": (c_int, []),
        "msb_sm_count": (c_int, []),
        "msb_last_error": (ctypes.c_char_p, []),
        "msb_project_point_fwd": (I, [P, P, P, I, I, I, F, F, P, P, V]),
        "msb_project_point_bwd": (I, [P, P, P, P, P, P, I, P, P, P, V]),
        "msb_compute_cov3d_fwd": (I, [P, P, P, I, P, V]),
        "msb_compute_cov3d_bwd": (I, [P, P, P, P, I, P, P, V]),
        "msb_ewa_project_fwd": (I, [P, P, P, P, P, P, I, I, I, P, P, P, V]),
        "msb_ewa_project_bwd": (I, [P, P, P, P, P, P, I, P, P, P, P, V]),
        "msb_preprocess_fwd": (I, [P, P, P, P, P, I, I, I, F, F, P, P, P, P, P, V]),
        "msb_preprocess_bwd": (I, [P, P, P, P, P, P, P, P, P, P, I, P, P, P, P, P, V]),
        "msb_compute_sh_fwd": (I, [P, P, P, I, I, I, P, V]),
        "msb_compute_sh_bwd": (I, [P, P, P, P, I, I, I, P, P, V]),
        "msb_sort_scan_workspace_bytes": (SZ, [I]),
        "msb_sort_scan": (I, [P, I, P, P, P, SZ, V]),
        "msb_sort_num_passes": (I, [I, I]),
        "msb_sort_workspace_bytes": (SZ, [I, LL, I, I]),
        "msb_sort_gaussian": (I, [P, P, P, P, I, LL, I, I, P, P, P, SZ, I, V]),
        "msb_blend_cpad": (I, [I]),
        "msb_blend_fwd_workspace_bytes": (SZ, [I, I]),
        "msb_blend_bwd_workspace_bytes": (SZ, [I, I]),
        "msb_alpha_blending_fwd": (I, [P, P, P, P, P, P, F, I, I, I, I, P, P, P, P, SZ, V]),
        "msb_alpha_blending_bwd": (I, [P, P, P, F, I, I, I, I, P, P, P, P, P, P, P, P, P, SZ, V]),
        "msb_blend_packed_fwd": (I, [P, P, P, P, F, I, I, I, P, P, P, V]),
        "msb_blend_packed_bwd": (I, [P, P, P, P, F, I, I, I, I, P, P, P, P, P, I, V]),
        "msb_render_preprocess_fwd": (I, [P] * 7 + [I, I, I, I, I, I, F, F, F, I] + [P] * 8 + [V]),
        "msb_render_preprocess_bwd": (I, [P] * 9 + [I, I, I, I, F, I, I] + [P] * 7 + [V]),
        # Level-1 drop-ins for msplat._C (integration/_C.py)
        "msb_compute_gaussian_key": (I, [P, P, P, P, I, LL, I, I, P, P, V]),
        "msb_compute_tile_gaussian_range": (I, [P, LL, I, I, P, V]),
        "msb_blend_pack": (I, [P, P, P, P, I, I, P, SZ, V]),
        "msb_adam_step": (I, [I, P, P, P, P, P, F, F, F, F, I, V]),
        # view batches
        "msb_sort_num_passes_views": (I, [I, I, I]),
        "msb_sort_workspace_bytes_views": (SZ, [I, I, LL, I, I]),
        "msb_sort_gaussian_views": (I, [P, P, P, P, I, I, LL, I, I, P, P, P, SZ, I, V]),
        "msb_blend_packed_fwd_views": (I, [P, P, P, P, F, I, I, I, I, P, P, P, V]),
        "msb_blend_packed_bwd_views": (I, [P, P, P, P, F, I, I, I, I, I, P, P, P, P, P, I, V]),
        "msb_blend_packed_count": (I, [P, P, P, I, I, I, P, V]),
        "msb_render_preprocess_fwd_views": (I, [P] * 7 + [I, I, I, LL, I, I, I, I, I, F, F, F, I] + [P] * 7 + [V]),
        "msb_render_preprocess_bwd_views": (I, [P] * 6 + [I] + [P] * 3 + [I, I, LL, I, I, I, F, I, I] + [P] * 7 + [V]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return sig


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"msplat_b200: {LIB_PATH} is missing. Build the CUDA library first "
                        "(python -m msplat_b200.build); there is no CPU or PyTorch fallback."
                    )
                L = ctypes.CDLL(LIB_PATH)
                _declare(L)
                _lib = L
    return _lib


def exported_symbols():
    """Names declared in include/msplat_b200.h that the library must export."""
    return sorted(_declare(lib()).keys())


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().msb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"msplat_b200.{what} failed (code {rc}): {msg}")


# When set to a list, every C-ABI call is bracketed by CUDA events on the launching stream and
# (name, start_event, end_event) is appended: bench.py uses it for per-kernel roofline numbers.
TIMING = None


def _invoke(what, fn, idx, args):
    st = c_void_p(torch._C._cuda_getCurrentRawStream(idx))
    if TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args, st)
        e1.record()
        TIMING.append((what, e0, e1))
        return rc
    return fn(*args, st)


def call(what: str, nlaunch: int, fn, device, *args):
    """Invoke one C-ABI entry point on `device`'s current stream; raise on a non-zero status.  The kernels run
    on the caller's current device: the device guard is only entered when `device` is not already current (the
    guard and the Stream object are the bulk of the per-call host cost otherwise)."""
    global _launches
    idx = device.index if isinstance(device, torch.device) else torch.device(device).index
    cur = torch.cuda.current_device()
    if idx is None or idx == cur:
        rc = _invoke(what, fn, cur if idx is None else idx, args)
    else:
        with torch.cuda.device(idx):
            rc = _invoke(what, fn, idx, args)
    if rc != 0:
        check(rc, what)
    with _count_lock:
        _launches += nlaunch


def count_launches(n: int):
    global _launches
    with _count_lock:
        _launches += n


def launches() -> int:
    return _launches


def reset_launches():
    global _launches
    _launches = 0


# ------------------------------------------------------------------------------------------------
# tensor plumbing
# ------------------------------------------------------------------------------------------------

def stream_ptr(device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t) -> c_void_p:
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def as_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """contiguous float32 CUDA tensor, 16-byte aligned (the reference calls .contiguous() too)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")  # reference: CHECK_CUDA, include/utils.h:9-10
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def as_i32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.int32:
        raise RuntimeError(f"{name} must be int32, got {t.dtype}")
    return t.contiguous()


def as_mask(t: torch.Tensor, name: str, n: int) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.bool:
        raise RuntimeError(f"{name} must be bool, got {t.dtype}")
    t = t.contiguous()
    if t.numel() != n:
        raise RuntimeError(f"{name} must have {n} elements, got {t.numel()}")
    return t


_pinned = threading.local()


def pinned_i64(device, n: int = 1) -> torch.Tensor:
    """A per-thread, per-device pinned int64[n] used for the sort stage's M read-back(s)."""
    cache = getattr(_pinned, "cache", None)
    if cache is None:
        cache = _pinned.cache = {}
    key = torch.device(device).index
    if key not in cache or cache[key].numel() < n:
        cache[key] = torch.zeros(max(n, 8), dtype=torch.int64).pin_memory()
    return cache[key]


_sm_count = {}


def sm_count(device) -> int:
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _sm_count:
        _sm_count[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _sm_count[idx]
