"""Drop-in for ``msplat._C`` (the reference's pybind11 module, /root/reference/msplat/src/ext.cpp:14-25)
over libmsplat_b200.so: the binding a maintainer of the reference adds to switch its CUDA backend to the
B200-native kernels without touching ``msplat/*.py`` ("Level 1" of INTEGRATION.md).

Every function takes and returns ``torch.Tensor`` exactly like its pybind11 original (prototypes:
/root/reference/msplat/include/*.h); outputs are allocated with torch (the library never allocates),
raw device pointers + sizes + the current CUDA stream go through ctypes.  Errors raise RuntimeError
(reference: TORCH_CHECK in include/utils.h:9-10).
"""
from __future__ import annotations

import torch

from msplat_b200 import _lib
from msplat_b200._lib import as_f32, as_i32, ptr

__all__ = [
    "project_point_forward", "project_point_backward", "compute_cov3d_forward", "compute_cov3d_backward",
    "ewa_project_forward", "ewa_project_backward", "compute_gaussian_key", "compute_tile_gaussian_range",
    "compute_sh_forward", "compute_sh_backward", "alpha_blending_forward", "alpha_blending_backward",
]

f32, i32, i64 = torch.float32, torch.int32, torch.int64


def _vis(visible, n):
    """torch.bool [P] or [P,1] -> contiguous one byte per Gaussian"""
    if visible is None:
        return None
    return _lib.as_mask(visible.reshape(-1), "visible", n)


# ---- project_point (include/project_point.h; src/project_point.cu:147-228) --------------------------------
def project_point_forward(xyz, intr, extr, W, H, nearest, extent):
    x, i, e = as_f32(xyz, "xyz"), as_f32(intr, "intr"), as_f32(extr, "extr")
    P = x.shape[0]
    uv = torch.empty((P, 2), dtype=f32, device=x.device)
    depth = torch.empty((P, 1), dtype=f32, device=x.device)
    _lib.call("project_point_forward", 1 if P else 0, _lib.lib().msb_project_point_fwd, x.device, ptr(x), ptr(i), ptr(e),
              P, int(W), int(H), float(nearest), float(extent), ptr(uv), ptr(depth))
    return uv, depth


def project_point_backward(xyz, intr, extr, W, H, uv, depth, dL_duv, dL_ddepth):
    x, i, e = as_f32(xyz, "xyz"), as_f32(intr, "intr"), as_f32(extr, "extr")
    d, guv, gd = as_f32(depth, "depth"), as_f32(dL_duv, "dL_duv"), as_f32(dL_ddepth, "dL_ddepth")
    P = x.shape[0]
    dxyz = torch.empty((P, 3), dtype=f32, device=x.device)
    dintr = torch.zeros((4,), dtype=f32, device=x.device)
    dextr = torch.zeros(tuple(extr.shape), dtype=f32, device=x.device)  # the reference returns [3,4]
    _lib.call("project_point_backward", 1 if P else 0, _lib.lib().msb_project_point_bwd, x.device, ptr(x), ptr(i), ptr(e),
              ptr(d), ptr(guv), ptr(gd), P, ptr(dxyz), ptr(dintr), ptr(dextr))
    return dxyz, dintr, dextr


# ---- compute_cov3d (include/compute_cov3d.h; src/compute_cov3d.cu:149-200) --------------------------------
def compute_cov3d_forward(scales, uquats, visible):
    s, q = as_f32(scales, "scales"), as_f32(uquats, "uquats")
    P = s.shape[0]
    cov = torch.empty((P, 6), dtype=f32, device=s.device)
    _lib.call("compute_cov3d_forward", 1 if P else 0, _lib.lib().msb_compute_cov3d_fwd, s.device, ptr(s), ptr(q),
              ptr(_vis(visible, P)), P, ptr(cov))
    return cov


def compute_cov3d_backward(scales, uquats, visible, dL_dcov3Ds):
    s, q, g = as_f32(scales, "scales"), as_f32(uquats, "uquats"), as_f32(dL_dcov3Ds, "dL_dcov3Ds")
    P = s.shape[0]
    ds = torch.empty((P, 3), dtype=f32, device=s.device)
    dq = torch.empty((P, 4), dtype=f32, device=s.device)
    _lib.call("compute_cov3d_backward", 1 if P else 0, _lib.lib().msb_compute_cov3d_bwd, s.device, ptr(s), ptr(q),
              ptr(_vis(visible, P)), ptr(g), P, ptr(ds), ptr(dq))
    return ds, dq


# ---- ewa_project (include/ewa_project.h; src/ewa_project.cu:254-345) --------------------------------------
def ewa_project_forward(xyz, cov3d, intr, extr, uv, W, H, visible):
    x, c, i, e, u = as_f32(xyz, "xyz"), as_f32(cov3d, "cov3d"), as_f32(intr, "intr"), as_f32(extr, "extr"), as_f32(uv, "uv")
    P = x.shape[0]
    conic = torch.empty((P, 3), dtype=f32, device=x.device)
    radius = torch.empty((P,), dtype=i32, device=x.device)
    tiles = torch.empty((P,), dtype=i32, device=x.device)
    _lib.call("ewa_project_forward", 1 if P else 0, _lib.lib().msb_ewa_project_fwd, x.device, ptr(x), ptr(c), ptr(i),
              ptr(e), ptr(u), ptr(_vis(visible, P)), P, int(W), int(H), ptr(conic), ptr(radius), ptr(tiles))
    return conic, radius, tiles


def ewa_project_backward(xyz, cov3d, intr, extr, radius, dL_dconic):
    x, c, i, e = as_f32(xyz, "xyz"), as_f32(cov3d, "cov3d"), as_f32(intr, "intr"), as_f32(extr, "extr")
    r, g = as_i32(radius, "radius"), as_f32(dL_dconic, "dL_dconic")
    P = x.shape[0]
    dxyz = torch.empty((P, 3), dtype=f32, device=x.device)
    dcov = torch.empty((P, 6), dtype=f32, device=x.device)
    dintr = torch.zeros((4,), dtype=f32, device=x.device)
    dextr = torch.zeros(tuple(extr.shape), dtype=f32, device=x.device)
    _lib.call("ewa_project_backward", 1 if P else 0, _lib.lib().msb_ewa_project_bwd, x.device, ptr(x), ptr(c), ptr(i),
              ptr(e), ptr(r), ptr(g), P, ptr(dxyz), ptr(dcov), ptr(dintr), ptr(dextr))
    return dxyz, dcov, dintr, dextr


# ---- sort stage (include/sort_gaussian.h; src/sort_gaussian.cu:74-142) ------------------------------------
def compute_gaussian_key(uv, depth, W, H, radius, tiles):
    """`tiles` is the inclusive int32 cumsum of tiles_touched (msplat/sort_gaussian.py:42)."""
    u, d = as_f32(uv, "uv"), as_f32(depth, "depth")
    r, t = as_i32(radius, "radius"), as_i32(tiles, "tiles")
    P = u.shape[0]
    M = int(t.reshape(-1)[P - 1].item()) if P > 0 else 0  # the reference's host sync (src/sort_gaussian.cu:91)
    key = torch.empty((M,), dtype=i64, device=u.device)
    idx = torch.empty((M,), dtype=i32, device=u.device)
    _lib.call("compute_gaussian_key", 1 if M else 0, _lib.lib().msb_compute_gaussian_key, u.device, ptr(u), ptr(d),
              ptr(r), ptr(t), P, M, int(W), int(H), ptr(key), ptr(idx))
    return key, idx


def compute_tile_gaussian_range(W, H, tiles, key_sorted):
    t = as_i32(tiles, "tiles")
    if key_sorted.dtype != i64 or not key_sorted.is_cuda:
        raise RuntimeError("key_sorted must be a CUDA int64 tensor")
    k = key_sorted.contiguous()
    T = ((int(W) + 15) // 16) * ((int(H) + 15) // 16)
    tr = torch.empty((T, 2), dtype=i32, device=t.device)
    _lib.call("compute_tile_gaussian_range", 1 if k.numel() else 0, _lib.lib().msb_compute_tile_gaussian_range, t.device,
              ptr(k), int(k.numel()), int(W), int(H), ptr(tr))
    return tr


# ---- compute_sh (include/compute_sh.h; src/compute_sh.cu:1696-1754) ---------------------------------------
def compute_sh_forward(shs, view_dirs, visible):
    s, d = as_f32(shs, "shs"), as_f32(view_dirs, "view_dirs")
    P, Cs, D = s.shape
    val = torch.empty((P, Cs), dtype=f32, device=s.device)
    _lib.call("compute_sh_forward", 1 if P else 0, _lib.lib().msb_compute_sh_fwd, s.device, ptr(s), ptr(d),
              ptr(_vis(visible, P)), P, Cs, D, ptr(val))
    return val


def compute_sh_backward(shs, view_dirs, visible, dL_dvalue):
    s, d, g = as_f32(shs, "shs"), as_f32(view_dirs, "view_dirs"), as_f32(dL_dvalue, "dL_dvalue")
    P, Cs, D = s.shape
    dshs = torch.empty_like(s)
    ddirs = torch.empty((P, 3), dtype=f32, device=s.device)
    _lib.call("compute_sh_backward", 1 if P else 0, _lib.lib().msb_compute_sh_bwd, s.device, ptr(s), ptr(d),
              ptr(_vis(visible, P)), ptr(g), P, Cs, D, ptr(dshs), ptr(ddirs))
    return dshs, ddirs


# ---- alpha_blending (include/alpha_blending.h; src/alpha_blending.cu:248-573) ------------------------------
def alpha_blending_forward(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H):
    u, c, o, f = as_f32(uv, "uv"), as_f32(conic, "conic"), as_f32(opacity, "opacity"), as_f32(feature, "feature")
    ids, tr = as_i32(idx_sorted, "idx_sorted"), as_i32(tile_range, "tile_range")
    P, C = f.shape
    L = _lib.lib()
    W, H = int(W), int(H)
    image = torch.empty((C, H, W), dtype=f32, device=f.device)
    final_T = torch.empty((H, W), dtype=f32, device=f.device)
    ncontrib = torch.empty((H, W), dtype=i32, device=f.device)
    packed = torch.empty((L.msb_blend_fwd_workspace_bytes(P, C),), dtype=torch.uint8, device=f.device)
    _lib.call("alpha_blending_forward", 2, L.msb_alpha_blending_fwd, f.device, ptr(u), ptr(c), ptr(o), ptr(f), ptr(ids),
              ptr(tr), float(bg), P, C, W, H, ptr(image), ptr(final_T), ptr(ncontrib), ptr(packed), packed.numel())
    return image, final_T, ncontrib


def alpha_blending_backward(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, final_T, ncontrib,
                            dL_drendered):
    u, c, o, f = as_f32(uv, "uv"), as_f32(conic, "conic"), as_f32(opacity, "opacity"), as_f32(feature, "feature")
    ids, tr = as_i32(idx_sorted, "idx_sorted"), as_i32(tile_range, "tile_range")
    g, fT, nc = as_f32(dL_drendered, "dL_drendered"), as_f32(final_T, "final_T"), as_i32(ncontrib, "ncontrib")
    P, C = f.shape
    L = _lib.lib()
    W, H = int(W), int(H)
    dev = f.device
    duv = torch.empty((P, 2), dtype=f32, device=dev)
    dconic = torch.empty((P, 3), dtype=f32, device=dev)
    dop = torch.empty((P, 1), dtype=f32, device=dev)
    dfeat = torch.empty((P, C), dtype=f32, device=dev)
    if P == 0 or C == 0:
        for t in (duv, dconic, dop, dfeat):
            t.zero_()
        return duv, dconic, dop, dfeat
    # the pybind11 signature hands the inputs over again instead of the forward's workspace: re-pack them
    packed = torch.empty((L.msb_blend_fwd_workspace_bytes(P, C),), dtype=torch.uint8, device=dev)
    _lib.call("blend_pack", 1, L.msb_blend_pack, dev, ptr(u), ptr(c), ptr(o), ptr(f), P, C, ptr(packed), packed.numel())
    ws = torch.empty((L.msb_blend_bwd_workspace_bytes(P, C),), dtype=torch.uint8, device=dev)
    _lib.call("alpha_blending_backward", 2, L.msb_alpha_blending_bwd, dev, ptr(f), ptr(ids), ptr(tr), float(bg), P, C, W,
              H, ptr(fT), ptr(nc), ptr(g), ptr(packed), ptr(duv), ptr(dconic), ptr(dop), ptr(dfeat), ptr(ws), ws.numel())
    return duv, dconic, dop, dfeat
