"""CPU restatement of the six msplat steps + rasterization().  TEST INFRASTRUCTURE ONLY.

Per-Gaussian stages are vectorised PyTorch (any float dtype; run on CPU), differentiable
through torch autograd exactly like the restatements inside the reference's own tests.
The sort stage is numpy integer arithmetic; the blend stage calls the C restatement
``oracle/blend_ref.c`` (and has a pure-Python loop twin for tiny cases).

Reference anchors (paths under /root/reference):
  project_point : msplat/src/project_point.cu:13-57,  test/test_project_points.py:8-52
  compute_cov3d : msplat/src/compute_cov3d.cu:14-58,  test/test_compute_cov3d.py:7-36
  ewa_project   : msplat/src/ewa_project.cu:16-83, msplat/include/utils.h:17-37,
                  test/test_ewa_project.py:11-107
  compute_sh    : msplat/src/compute_sh.cu:116-503,1600-1637, test/test_compute_sh.py:8-13
  sort_gaussian : msplat/sort_gaussian.py:42-52, msplat/src/sort_gaussian.cu:17-71
  alpha_blending: msplat/src/alpha_blending.cu:16-246, test/test_alpha_blending.py:6-63
  rasterization : msplat/__init__.py:70-93
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional, Tuple

import numpy as np
import torch

TILE = 16  # msplat/include/config.h:7-8

# --------------------------------------------------------------------------------------
# project_point
# --------------------------------------------------------------------------------------


def project_point(xyz, intr, extr, W, H, nearest: float = 0.0, extent: float = 1.3):
    """uv [P,2], depth [P,1].  Follows project_point.cu:27-56.

    ``t = R p + T``; ``uv = f * t.xy / (t.z + 1e-7) + c - 0.5`` (the reciprocal and the
    final ``- 0.5`` are evaluated in double in the reference, :31,34-35); near cull only
    if ``nearest > 0`` (:39-41); extent cull (:43-51); culled rows are all-zero (:52-53
    with the pre-zeroed outputs :161-162).
    """
    dt = xyz.dtype
    R = extr[:3, :3]
    T = extr[:3, 3]
    t = xyz @ R.T + T
    z = t[:, 2]
    norm = (1.0 / (z.double() + 1e-7))
    u = (intr[0] * t[:, 0]).double() * norm + intr[2].double() - 0.5
    v = (intr[1] * t[:, 1]).double() * norm + intr[3].double() - 0.5
    u = u.to(dt)
    v = v.to(dt)
    cull = torch.zeros_like(z, dtype=torch.bool)
    if nearest > 0:
        cull = cull | (z <= nearest)
    if extent > 0:
        xmin, xmax = (1 - extent) * W * 0.5, (1 + extent) * W * 0.5
        ymin, ymax = (1 - extent) * H * 0.5, (1 + extent) * H * 0.5
        cull = cull | (u < xmin) | (u > xmax) | (v < ymin) | (v > ymax)
    keep = (~cull).to(dt)
    uv = torch.stack([u, v], dim=-1)
    # nan (t.z + 1e-7 == 0 etc.) never compares inside the limits in the reference either
    uv = torch.where(cull[:, None], torch.zeros_like(uv), uv)
    depth = torch.where(cull, torch.zeros_like(z), z)
    del keep
    return uv, depth[:, None]


# --------------------------------------------------------------------------------------
# compute_cov3d
# --------------------------------------------------------------------------------------


def _quat_to_rot(q):
    """(r, x, y, z) -> standard rotation matrix, *not* normalised (compute_cov3d.cu:24-39)."""
    r, x, y, z = q.unbind(-1)
    R = torch.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y),
        ],
        dim=-1,
    )
    return R.reshape(*q.shape[:-1], 3, 3)


def compute_cov3d(scales, uquats, visible=None):
    """cov3d [P,6] = upper triangle of R S^2 R^T (compute_cov3d.cu:41-58); rows with
    ``visible == False`` stay zero (:126 with the pre-zeroed output)."""
    R = _quat_to_rot(uquats)
    M = R * scales[:, None, :]  # R @ diag(s)
    S = M @ M.transpose(1, 2)
    out = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=-1)
    if visible is not None:
        out = out * visible.reshape(-1, 1).to(out.dtype)
    return out


# --------------------------------------------------------------------------------------
# ewa_project
# --------------------------------------------------------------------------------------


def _f2i_trunc(x: torch.Tensor) -> torch.Tensor:
    """float -> int32 like CUDA's cvt.rzi.s32.f32: truncate toward zero, saturate, NaN -> 0."""
    x = torch.nan_to_num(x.detach().double(), nan=0.0, posinf=2.0**31 - 1, neginf=-(2.0**31))
    return torch.clamp(torch.trunc(x), -(2.0**31), 2.0**31 - 1).to(torch.int64)


def get_rect(uv, radius, W, H):
    """Tile rectangle [min, max) of a splat, include/utils.h:17-37.  Returns int64 tensors
    (xmin, ymin, xmax, ymax).  ``radius`` is the *integer* radius (the reference passes the
    float radius through an int parameter)."""
    gx = (W + TILE - 1) // TILE
    gy = (H + TILE - 1) // TILE
    r = radius.to(uv.dtype).reshape(-1)
    x, y = uv[:, 0].detach(), uv[:, 1].detach()
    xmin = torch.clamp(_f2i_trunc((x - r) / TILE), 0, gx)
    ymin = torch.clamp(_f2i_trunc((y - r) / TILE), 0, gy)
    xmax = torch.clamp(_f2i_trunc((x + r + (TILE - 1)) / TILE), 0, gx)
    ymax = torch.clamp(_f2i_trunc((y + r + (TILE - 1)) / TILE), 0, gy)
    return xmin, ymin, xmax, ymax


def ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible=None):
    """conic [P,3] float, radius [P] int32, tiles [P] int32.  Follows ewa_project.cu:27-82.

    J, W, T = J W, cov2D = T Sigma T^T (+0.3 low-pass on the diagonal, :57-59); ``det == 0``
    -> skipped (:62); radius = ceil(3 sqrt(max eigenvalue)) with ``max(0.1, mid^2-det)``
    (:65-68); tile rectangle via get_rect, zero area -> skipped (:72-74);
    conic = inverse of cov2D (:76-79).  Skipped rows are all-zero.
    """
    P = xyz.shape[0]
    dt = xyz.dtype
    if visible is None:
        visible = torch.ones(P, dtype=torch.bool, device=xyz.device)
    visible = visible.reshape(-1).bool()
    fx, fy = intr[0], intr[1]
    Wm = extr[:3, :3]
    t = xyz @ Wm.T + extr[:3, 3]
    tx, ty, tz = t.unbind(-1)
    zero = torch.zeros_like(tz)
    # J rows (the third row of the reference's 3x3 J is zero and never used downstream)
    J = torch.stack(
        [
            torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz)], dim=-1),
            torch.stack([zero, fy / tz, -(fy * ty) / (tz * tz)], dim=-1),
        ],
        dim=-2,
    )  # [P,2,3]
    T = J @ Wm  # [P,2,3]
    c = cov3d
    V = torch.stack(
        [
            torch.stack([c[:, 0], c[:, 1], c[:, 2]], dim=-1),
            torch.stack([c[:, 1], c[:, 3], c[:, 4]], dim=-1),
            torch.stack([c[:, 2], c[:, 4], c[:, 5]], dim=-1),
        ],
        dim=-2,
    )
    cov2d = T @ V @ T.transpose(1, 2)
    a = cov2d[:, 0, 0] + 0.3
    b = cov2d[:, 0, 1]
    d = cov2d[:, 1, 1] + 0.3
    det = a * d - b * b
    ok = visible & (det != 0)
    mid = 0.5 * (a + d)
    disc = torch.clamp(mid * mid - det, min=0.1)
    lam = torch.maximum(mid + torch.sqrt(disc), mid - torch.sqrt(disc))
    radius_f = torch.ceil(3.0 * torch.sqrt(lam))
    radius_i = _f2i_trunc(radius_f)
    xmin, ymin, xmax, ymax = get_rect(uv, radius_i, W, H)
    tiles = (xmax - xmin) * (ymax - ymin)
    ok = ok & (tiles != 0)
    okf = ok.to(dt)
    safe_det = torch.where(ok, det, torch.ones_like(det))
    conic = torch.stack([d / safe_det, -b / safe_det, a / safe_det], dim=-1) * okf[:, None]
    radius = torch.where(ok, radius_i, torch.zeros_like(radius_i)).to(torch.int32)
    tiles = torch.where(ok, tiles, torch.zeros_like(tiles)).to(torch.int32)
    return conic, radius, tiles


# --------------------------------------------------------------------------------------
# compute_sh
# --------------------------------------------------------------------------------------

_SH_CACHE = {}


def _sh_tables(deg_max: int = 10):
    """Per-(l,m) data for the real SH basis in the polynomial form the reference uses:
    ``Y_l^m = N_lm * Q_lm(z) * {Re, Im}((x + i y)^|m|)`` with ``Q_lm = d^|m|/dz^|m| P_l(z)``
    and ``N_lm = (-1)^m sqrt(2) K_l^|m|`` (``K_l^0`` for m = 0).  This identity was checked
    against every expression in compute_sh.cu:116-503 off the unit sphere (see
    tests/golden/make_golden.py); degrees 2 and 3 contain the homogeneous exceptions
    handled in :func:`sh_basis`.
    """
    if deg_max in _SH_CACHE:
        return _SH_CACHE[deg_max]
    from numpy.polynomial import legendre as L
    from numpy.polynomial import polynomial as Pn

    tab = {}
    for l in range(deg_max + 1):
        pl = L.leg2poly([0.0] * l + [1.0])  # ascending power coefficients of P_l(z)
        for m in range(0, l + 1):
            q = Pn.polyder(pl, m) if m > 0 else pl
            K = math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - m) / math.factorial(l + m))
            N = K if m == 0 else ((-1) ** m) * math.sqrt(2.0) * K
            tab[(l, m)] = (N, np.array(q, dtype=np.float64))
    _SH_CACHE[deg_max] = tab
    return tab


def _polyval(coef, z):
    out = torch.zeros_like(z) + float(coef[-1])
    for c in coef[-2::-1]:
        out = out * z + float(c)
    return out


def sh_basis(dirs: torch.Tensor, D: int) -> torch.Tensor:
    """Basis values [..., D] in the reference's ordering (index l^2 + l + m, negative m =
    imaginary part; compute_sh.cu:125-127 for the degree-1 signs)."""
    deg = int(round(math.sqrt(D))) - 1
    assert (deg + 1) ** 2 == D and 0 <= deg <= 10, "D must be (deg+1)^2 with deg <= 10"
    tab = _sh_tables(10)
    x, y, z = dirs.unbind(-1)
    # Re/Im of (x + i y)^m
    re = [torch.ones_like(x)]
    im = [torch.zeros_like(x)]
    for m in range(1, deg + 1):
        re.append(re[-1] * x - im[-1] * y)
        im.append(re[-2] * y + im[-1] * x)
    out = [None] * D
    x2, y2, z2 = x * x, y * y, z * z
    for l in range(deg + 1):
        for m in range(-l, l + 1):
            N, q = tab[(l, abs(m))]
            ang = re[m] if m >= 0 else im[-m]
            val = N * _polyval(q, z) * ang
            # homogeneous exceptions: compute_sh.cu:141 (l=2,m=0), :157-159 (l=3, m=-1,0,1)
            if l == 2 and m == 0:
                val = N * 0.5 * (2.0 * z2 - x2 - y2)
            elif l == 3 and m == 0:
                val = N * 0.5 * z * (2.0 * z2 - 3.0 * x2 - 3.0 * y2)
            elif l == 3 and abs(m) == 1:
                val = N * 1.5 * (4.0 * z2 - x2 - y2) * ang
            out[l * l + l + m] = val
    return torch.stack(out, dim=-1)


def compute_sh(shs, view_dirs, visible=None):
    """value [P,Cs] = sum_d basis_d(dir) * shs[:, :, d]; no +0.5, no clamp
    (compute_sh.cu:1611-1636); invisible rows are zero (:1608)."""
    P, Cs, D = shs.shape
    B = sh_basis(view_dirs, D)
    val = (B[:, None, :] * shs).sum(dim=-1)
    if visible is not None:
        val = val * visible.reshape(-1, 1).to(val.dtype)
    return val


# --------------------------------------------------------------------------------------
# sort_gaussian (integer stage, numpy)
# --------------------------------------------------------------------------------------


def sort_gaussian(uv, depth, W, H, radius, tiles) -> Tuple[torch.Tensor, torch.Tensor]:
    """idx_sorted int32 [M], tile_range int32 [T,2].

    sort_gaussian.py:42 inclusive int32 cumsum of ``tiles``; sort_gaussian.cu:26-42 each
    Gaussian with ``radius > 0`` writes ``key = (tile_id << 32) | depth_bits`` and its index
    for every tile of its rectangle (rows, then columns) starting at ``cumsum[idx-1]``; slots
    never written stay ``(key 0, idx 0)`` (pre-zeroed :98-99); sort_gaussian.py:49-50 sort by
    key (ties keep slot order -- a stable sort, as torch's CUDA radix sort is);
    sort_gaussian.cu:45-71 tile_range from boundaries of ``key >> 32`` (empty tiles (0,0)).

    Defined-behaviour choices where the reference is UB: depth bits are masked to 32 bits
    (SURVEY H4); a Gaussian writes at most ``tiles[idx]`` entries (never into a neighbour's
    slots).
    """
    uv_n = uv.detach().cpu().float().numpy().reshape(-1, 2)
    depth_n = depth.detach().cpu().float().numpy().reshape(-1)
    rad = radius.detach().cpu().numpy().reshape(-1).astype(np.int64)
    til = tiles.detach().cpu().numpy().reshape(-1).astype(np.int32)
    P = uv_n.shape[0]
    gx = (W + TILE - 1) // TILE
    gy = (H + TILE - 1) // TILE
    T = gx * gy
    tile_range = np.zeros((T, 2), dtype=np.int32)
    if P == 0:
        return torch.zeros(0, dtype=torch.int32), torch.from_numpy(tile_range)
    csum = np.cumsum(til, dtype=np.int32).astype(np.int64)
    M = int(csum[-1])
    start = np.concatenate([[0], csum[:-1]])
    xmin, ymin, xmax, ymax = (
        t.numpy() for t in get_rect(torch.from_numpy(uv_n), torch.from_numpy(rad), W, H)
    )
    area = (xmax - xmin) * (ymax - ymin)
    n = np.where(rad > 0, np.minimum(area, til.astype(np.int64)), 0)
    n = np.maximum(n, 0)
    keys = np.zeros(M, dtype=np.int64)
    ids = np.zeros(M, dtype=np.int32)
    tot = int(n.sum())
    if tot > 0:
        gid = np.repeat(np.arange(P, dtype=np.int64), n)
        first = np.repeat(np.cumsum(n) - n, n)
        local = np.arange(tot, dtype=np.int64) - first
        w = np.maximum(xmax - xmin, 1)[gid]
        ty = ymin[gid] + local // w
        tx = xmin[gid] + local % w
        dbits = depth_n.view(np.uint32).astype(np.int64)[gid]
        pos = start[gid] + local
        keys[pos] = ((ty * gx + tx) << 32) | dbits
        ids[pos] = gid.astype(np.int32)
    order = np.argsort(keys, kind="stable")
    keys_s = keys[order]
    ids_s = ids[order]
    if M > 0:
        tid = (keys_s >> 32).astype(np.int64)
        tile_range[tid[0], 0] = 0
        tile_range[tid[-1], 1] = M
        cut = np.nonzero(tid[1:] != tid[:-1])[0] + 1
        tile_range[tid[cut - 1], 1] = cut
        tile_range[tid[cut], 0] = cut
    return torch.from_numpy(ids_s.copy()), torch.from_numpy(tile_range)


# --------------------------------------------------------------------------------------
# alpha_blending (C restatement behind ctypes + python loop twin)
# --------------------------------------------------------------------------------------

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_blend_ref(force: bool = False) -> str:
    """Compile oracle/blend_ref.c with gcc into oracle/_build/libblend_ref.so."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libblend_ref.so")
    src = os.path.join(_HERE, "blend_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src, "-o", so, "-lm"]
        subprocess.check_call(cmd)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_blend_ref())
        _LIB.blend_forward_ref.restype = ctypes.c_int64
        _LIB.blend_backward_ref.restype = None
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _np32(t, shape=None):
    a = np.ascontiguousarray(t.detach().cpu().float().numpy())
    return a.reshape(shape) if shape is not None else a


def alpha_blending_forward(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H):
    """(image [C,H,W], final_T [H,W], ncontrib [H,W], pairs) -- alpha_blending.cu:16-110."""
    P, C = feature.shape
    uv_n, co_n, op_n, fe_n = _np32(uv, (P, 2)), _np32(conic, (P, 3)), _np32(opacity, (P,)), _np32(feature, (P, C))
    ids = np.ascontiguousarray(idx_sorted.detach().cpu().numpy().astype(np.int32))
    tr = np.ascontiguousarray(tile_range.detach().cpu().numpy().astype(np.int32))
    img = np.zeros((C, H, W), dtype=np.float32)
    fT = np.zeros((H, W), dtype=np.float32)
    nc = np.zeros((H, W), dtype=np.int32)
    pairs = _lib().blend_forward_ref(
        ctypes.c_int(P), ctypes.c_int(C), ctypes.c_int(W), ctypes.c_int(H), _fp(uv_n), _fp(co_n), _fp(op_n),
        _fp(fe_n), _fp(ids), _fp(tr), ctypes.c_float(bg), _fp(img), _fp(fT), _fp(nc))
    return torch.from_numpy(img), torch.from_numpy(fT), torch.from_numpy(nc), int(pairs)


def alpha_blending_backward(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, final_T, ncontrib, dL_dimage):
    """(dL_duv [P,2], dL_dconic [P,3], dL_dopacity [P,1], dL_dfeature [P,C]) --
    alpha_blending.cu:112-246."""
    P, C = feature.shape
    uv_n, co_n, op_n, fe_n = _np32(uv, (P, 2)), _np32(conic, (P, 3)), _np32(opacity, (P,)), _np32(feature, (P, C))
    ids = np.ascontiguousarray(idx_sorted.detach().cpu().numpy().astype(np.int32))
    tr = np.ascontiguousarray(tile_range.detach().cpu().numpy().astype(np.int32))
    fT = _np32(final_T, (H, W))
    nc = np.ascontiguousarray(ncontrib.detach().cpu().numpy().astype(np.int32))
    dimg = _np32(dL_dimage, (C, H, W))
    duv = np.zeros((P, 2), np.float32)
    dco = np.zeros((P, 3), np.float32)
    dop = np.zeros((P, 1), np.float32)
    dfe = np.zeros((P, C), np.float32)
    _lib().blend_backward_ref(
        ctypes.c_int(P), ctypes.c_int(C), ctypes.c_int(W), ctypes.c_int(H), _fp(uv_n), _fp(co_n), _fp(op_n),
        _fp(fe_n), _fp(ids), _fp(tr), ctypes.c_float(bg), _fp(fT), _fp(nc), _fp(dimg), _fp(duv), _fp(dco),
        _fp(dop), _fp(dfe))
    return torch.from_numpy(duv), torch.from_numpy(dco), torch.from_numpy(dop), torch.from_numpy(dfe)


class _BlendFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H):
        img, fT, nc, _ = alpha_blending_forward(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H)
        ctx.save_for_backward(uv, conic, opacity, feature, idx_sorted, tile_range, fT, nc)
        ctx.meta = (bg, W, H)
        return img.to(feature.dtype)

    @staticmethod
    def backward(ctx, g):
        uv, conic, opacity, feature, ids, tr, fT, nc = ctx.saved_tensors
        bg, W, H = ctx.meta
        duv, dco, dop, dfe = alpha_blending_backward(uv, conic, opacity, feature, ids, tr, bg, W, H, fT, nc, g)
        return (duv.to(uv.dtype), dco.to(conic.dtype), dop.reshape(opacity.shape).to(opacity.dtype),
                dfe.to(feature.dtype), None, None, None, None, None)


def alpha_blending(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H):
    """Differentiable oracle blend (float32 arithmetic inside, like the reference)."""
    return _BlendFn.apply(uv, conic, opacity, feature, idx_sorted, tile_range, float(bg), int(W), int(H))


def alpha_blending_loop(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H):
    """Pure-Python per-pixel loop (torch scalars, autograd-able) for tiny shapes: the form
    the reference's own test uses (test/test_alpha_blending.py:6-63) but with the CUDA
    kernel's ``next_T < 1e-4`` termination (alpha_blending.cu:90) rather than the test's
    ``<=``, and with no gradient through the 0.99 clamp's *selection* only (python ``min``),
    which is what the test restatement does."""
    C = feature.shape[1]
    gx = (W + TILE - 1) // TILE
    rows = []
    for i in range(H):
        row = []
        for j in range(W):
            tile = (i // TILE) * gx + (j // TILE)
            T = torch.ones((), dtype=uv.dtype)
            acc = torch.zeros(C, dtype=uv.dtype)
            for e in range(int(tile_range[tile, 0]), int(tile_range[tile, 1])):
                g = int(idx_sorted[e])
                d = uv[g] - torch.tensor([j, i], dtype=uv.dtype)
                power = -0.5 * (conic[g, 0] * d[0] * d[0] + conic[g, 2] * d[1] * d[1]) - conic[g, 1] * d[0] * d[1]
                if power > 0:
                    continue
                a = opacity[g].reshape(()) * torch.exp(power)
                alpha = a if a < 0.99 else torch.full((), 0.99, dtype=uv.dtype)
                if alpha < 1.0 / 255.0:
                    continue
                nT = T * (1 - alpha)
                if nT < 1e-4:
                    break
                acc = acc + feature[g] * alpha * T
                T = nT
            row.append(acc + T * bg)
        rows.append(torch.stack(row))
    return torch.stack(rows).permute(2, 0, 1)


# --------------------------------------------------------------------------------------
# rasterization
# --------------------------------------------------------------------------------------


def rasterization(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg, ndc: Optional[torch.Tensor] = None):
    """msplat/__init__.py:70-93: project -> visible = depth != 0 -> cov3d -> ewa -> sort -> blend."""
    uv, depth = project_point(xyz, intr, extr, W, H)
    visible = depth != 0
    cov3d = compute_cov3d(scale, rotate, visible)
    conic, radius, tiles = ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
    ids, tr = sort_gaussian(uv, depth, W, H, radius, tiles)
    return alpha_blending(uv, conic, opacity, feature, ids, tr, bg, W, H)
