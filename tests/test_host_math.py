"""The product's per-Gaussian / per-pair math (msplat_b200/csrc/{geom,sh_eval,blend_math}.cuh),
compiled for the host by tools/host_check.cu, checked against the oracle on a CPU-only box.
(On the host the MUFU approximations are IEEE ops, so this is a tolerance-level check; the
bit-level parity against the reference CUDA build is asserted by the -m gpu tests.)
"""
import ctypes
import math

import numpy as np
import torch

import oracle
from conftest import fp

T = torch.from_numpy


def _camera(W, H):
    th = 0.3
    intr = torch.tensor([700.0, 710.0, W / 2, H / 2])
    extr = torch.tensor([[math.cos(th), 0, math.sin(th), 0.1], [0, 1, 0, -0.2], [-math.sin(th), 0, math.cos(th), 4.0]])
    return intr, extr


def _c(t):
    return np.ascontiguousarray(t.detach().numpy())


def test_project_point_fwd_bwd(host_check):
    torch.manual_seed(0)
    P, W, H = 4000, 800, 600
    xyz = torch.randn(P, 3) * 1.5
    intr, extr = _camera(W, H)
    uv_o, d_o = oracle.project_point(xyz, intr, extr, W, H)
    uv, dep = np.zeros((P, 2), np.float32), np.zeros(P, np.float32)
    host_check.hc_project_fwd(P, fp(_c(xyz)), fp(_c(intr)), fp(_c(extr)), W, H, ctypes.c_float(0.0),
                              ctypes.c_float(1.3), fp(uv), fp(dep))
    assert ((dep == 0) == (d_o.numpy()[:, 0] == 0)).all() and (dep == 0).sum() > 100
    np.testing.assert_allclose(uv, uv_o.numpy(), rtol=1e-5, atol=2e-4)
    np.testing.assert_array_equal(dep, d_o.numpy()[:, 0])
    # near-plane culling (nearest > 0) -- test/test_project_points.py uses 0.2
    uv_n, d_n = oracle.project_point(xyz, intr, extr, W, H, nearest=3.5)
    host_check.hc_project_fwd(P, fp(_c(xyz)), fp(_c(intr)), fp(_c(extr)), W, H, ctypes.c_float(3.5),
                              ctypes.c_float(1.3), fp(uv), fp(dep))
    assert ((dep == 0) == (d_n.numpy()[:, 0] == 0)).all()
    # backward incl. camera gradients, against float64 autograd of the oracle
    x64 = xyz.double().requires_grad_()
    i64, e64 = intr.double().requires_grad_(), extr.double().requires_grad_()
    uv_r, d_r = oracle.project_point(x64, i64, e64, W, H)
    guv, gd = torch.randn(P, 2), torch.randn(P)
    ((uv_r * guv.double()).sum() + (d_r[:, 0] * gd.double()).sum()).backward()
    dx, cam = np.zeros((P, 3), np.float32), np.zeros(16, np.float32)
    host_check.hc_project_fwd(P, fp(_c(xyz)), fp(_c(intr)), fp(_c(extr)), W, H, ctypes.c_float(0.0),
                              ctypes.c_float(1.3), fp(uv), fp(dep))
    host_check.hc_project_bwd(P, fp(_c(xyz)), fp(_c(intr)), fp(_c(extr)), fp(dep), fp(_c(guv)), fp(_c(gd)), fp(dx),
                              fp(cam))
    np.testing.assert_allclose(dx, x64.grad.numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(cam[:4], i64.grad.numpy(), rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(cam[4:], e64.grad.numpy().reshape(-1), rtol=1e-3, atol=0.5)


def test_cov3d_fwd_bwd(host_check):
    torch.manual_seed(1)
    P = 3000
    s = torch.rand(P, 3) + 0.05
    q = torch.randn(P, 4)  # deliberately NOT normalised
    cov = np.zeros((P, 6), np.float32)
    host_check.hc_cov3d_fwd(P, fp(_c(s)), fp(_c(q)), fp(cov))
    np.testing.assert_allclose(cov, oracle.compute_cov3d(s, q).numpy(), rtol=1e-5, atol=1e-5)
    s64, q64 = s.double().requires_grad_(), q.double().requires_grad_()
    g = torch.randn(P, 6)
    (oracle.compute_cov3d(s64, q64) * g.double()).sum().backward()
    ds, dq = np.zeros((P, 3), np.float32), np.zeros((P, 4), np.float32)
    host_check.hc_cov3d_bwd(P, fp(_c(s)), fp(_c(q)), fp(_c(g)), fp(ds), fp(dq))
    np.testing.assert_allclose(ds, s64.grad.numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(dq, q64.grad.numpy(), rtol=1e-4, atol=1e-3)


def test_ewa_fwd_bwd(host_check):
    torch.manual_seed(2)
    P, W, H = 5000, 800, 600
    xyz = torch.randn(P, 3) * 1.5
    intr, extr = _camera(W, H)
    uv, depth = oracle.project_point(xyz, intr, extr, W, H)
    vis = (depth != 0).reshape(-1)
    s = (torch.rand(P, 3) + 0.05) * 0.2
    q = torch.randn(P, 4)
    q = q / q.norm(dim=-1, keepdim=True)
    cov = oracle.compute_cov3d(s, q, vis)
    conic_o, rad_o, til_o = oracle.ewa_project(xyz, cov, intr, extr, uv, W, H, vis)
    conic, rad, til = np.zeros((P, 3), np.float32), np.zeros(P, np.int32), np.zeros(P, np.int32)
    host_check.hc_ewa_fwd(P, fp(_c(xyz)), fp(_c(cov)), fp(_c(intr)), fp(_c(extr)), fp(_c(uv)),
                          fp(_c(vis).astype(np.uint8)), W, H, fp(conic), fp(rad), fp(til))
    # integer outputs may differ only where 3*sqrt(lambda) sits within an ulp of an integer
    assert (rad != rad_o.numpy()).mean() < 2e-3 and (til != til_o.numpy()).mean() < 2e-3
    same = rad == rad_o.numpy()
    np.testing.assert_allclose(conic[same], conic_o.numpy()[same], rtol=2e-4, atol=1e-6)
    assert (til > 0).sum() > 1000
    x64, c64 = xyz.double().requires_grad_(), cov.double().requires_grad_()
    i64, e64 = intr.double().requires_grad_(), extr.double().requires_grad_()
    c_r, _, _ = oracle.ewa_project(x64, c64, i64, e64, uv.double(), W, H, vis)
    gc = torch.randn(P, 3)
    (c_r * gc.double()).sum().backward()
    dx, dcov, cam = np.zeros((P, 3), np.float32), np.zeros((P, 6), np.float32), np.zeros(16, np.float32)
    host_check.hc_ewa_bwd(P, fp(_c(xyz)), fp(_c(cov)), fp(_c(intr)), fp(_c(extr)), fp(_c(rad_o)), fp(_c(gc)), fp(dx),
                          fp(dcov), fp(cam))
    sx, sc = np.abs(x64.grad.numpy()).max(), np.abs(c64.grad.numpy()).max()
    assert np.abs(dx - x64.grad.numpy()).max() < 1e-4 * sx
    assert np.abs(dcov - c64.grad.numpy()).max() < 1e-4 * sc
    np.testing.assert_allclose(cam[:2], i64.grad.numpy()[:2], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(cam[4:], e64.grad.numpy().reshape(-1), rtol=2e-3, atol=1e-3)


def test_sh_basis_and_gradient(host_check, golden):
    g = golden("sh_basis.npz")
    dirs = g["dirs"][:64].astype(np.float32)  # unit directions
    for deg in range(11):
        D = (deg + 1) ** 2
        out = np.zeros((64, D), np.float32)
        assert host_check.hc_sh_basis(deg, 64, fp(dirs), fp(out)) == 0
        ref = oracle.sh_basis(T(dirs.astype(np.float64)), D).numpy()
        np.testing.assert_allclose(out, ref, rtol=1e-4, atol=5e-5 if deg <= 6 else 5e-4)
        d64 = T(dirs.astype(np.float64)).requires_grad_()
        w = torch.randn(64, D, dtype=torch.float64)
        ((oracle.sh_basis(d64, D) * w).sum() + 0.0 * d64.sum()).backward()  # deg 0 has no dir dependence
        gout = np.zeros((64, 3), np.float32)
        assert host_check.hc_sh_grad(deg, 64, fp(dirs), fp(_c(w.float())), fp(gout)) == 0
        scale = max(float(d64.grad.abs().max()), 1.0)
        assert np.abs(gout - d64.grad.numpy()).max() < 2e-4 * scale, deg
    # off the unit sphere the same polynomials must hold (dL_ddir is their derivative): relative check
    off = g["dirs"][64:].astype(np.float32)
    out = np.zeros((64, 121), np.float32)
    host_check.hc_sh_basis(10, 64, fp(off), fp(out))
    ref = g["basis_cuda_text"][64:]
    assert np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1.0)) < 2e-3


def test_blend_pair_math_and_cull_extent(host_check):
    """One pixel over a random list: product pair math == oracle C loop; and the culling box never
    excludes a pair that passes the alpha test."""
    rng = np.random.default_rng(5)
    n, C = 300, 4
    uv = (rng.random((n, 2)) * 16).astype(np.float32)
    A = rng.normal(size=(n, 2, 2)) * 0.4
    cv = A @ A.transpose(0, 2, 1) + 0.05 * np.eye(2)
    cinv = np.linalg.inv(cv)
    conic = np.stack([cinv[:, 0, 0], cinv[:, 0, 1], cinv[:, 1, 1]], -1).astype(np.float32)
    op = rng.random(n).astype(np.float32)
    op[:20] *= 0.0035  # below 1/255: can never pass the alpha test
    feat = rng.random((n, C)).astype(np.float32)
    F, last = np.zeros(C, np.float32), ctypes.c_int(0)
    host_check.hc_blend_pixel.restype = ctypes.c_float
    Tf = host_check.hc_blend_pixel(n, fp(uv), fp(conic), fp(op), fp(feat), C, ctypes.c_float(7.0), ctypes.c_float(5.0),
                                   fp(F), ctypes.byref(last))
    # oracle: a 16x16 image whose single tile lists all n Gaussians in order
    ids = torch.arange(n, dtype=torch.int32)
    tr = torch.tensor([[0, n]], dtype=torch.int32)
    img, fT, nc, _ = oracle.alpha_blending_forward(T(uv), T(conic), T(op), T(feat), ids, tr, 0.0, 16, 16)
    np.testing.assert_allclose(F, img[:, 5, 7].numpy(), rtol=1e-5, atol=1e-6)
    assert abs(Tf - float(fT[5, 7])) < 1e-6 and last.value == int(nc[5, 7])
    # culling extents are conservative
    ext = np.zeros((n, 4), np.float32)  # hx, hy, hs (|dx + dy|), ht (|dx - dy|) after the FP16 round trip
    host_check.hc_cull_extent(n, fp(conic), fp(op), fp(ext))
    ys, xs = np.mgrid[0:16, 0:16]
    culled_blocks = 0
    for j in range(n):
        dx, dy = uv[j, 0] - xs, uv[j, 1] - ys
        power = -0.5 * (conic[j, 0] * dx * dx + conic[j, 2] * dy * dy) - conic[j, 1] * dx * dy
        alpha = np.minimum(0.99, op[j] * np.exp(power))
        passes = (power <= 0) & (alpha >= 1.0 / 255.0)
        if passes.any():
            assert (np.abs(dx[passes]) <= ext[j, 0]).all() and (np.abs(dy[passes]) <= ext[j, 1]).all(), j
            assert (np.abs(dx[passes] + dy[passes]) <= ext[j, 2]).all(), j
            assert (np.abs(dx[passes] - dy[passes]) <= ext[j, 3]).all(), j
        # the kernels' block test never rejects an 8x4 block that holds a passing pixel
        for by in range(0, 16, 4):
            for bx in range(0, 16, 8):
                miss = host_check.hc_cull_miss(ctypes.c_float(uv[j, 0]), ctypes.c_float(uv[j, 1]), fp(conic[j]),
                                               ctypes.c_float(op[j]), ctypes.c_float(bx), ctypes.c_float(by),
                                               ctypes.c_float(8.0), ctypes.c_float(4.0))
                if miss:
                    culled_blocks += 1
                    assert not passes[by:by + 4, bx:bx + 8].any(), (j, bx, by)
    assert culled_blocks > 0
    assert (ext[:20] < 0).all()  # opacity < 1/255 can never contribute


def test_cull_extent_fp16_packing_rounds_up(host_check):
    """The culling extents travel as FP16: the round trip never shrinks a value (an extent that
    became smaller could cull a contributing pair), overflows become +inf (= no culling)."""
    rng = np.random.default_rng(9)
    vals = np.concatenate([np.exp(rng.uniform(np.log(1e-4), np.log(6e4), 4000)),
                           [0.0, 1e-8, 65504.0, 65505.0, 7e4, 1e9, np.inf, -np.inf]]).astype(np.float32)
    vals = np.resize(vals, (len(vals) + 3) // 4 * 4)
    out = np.zeros_like(vals)
    host_check.hc_cull_pack_roundtrip(len(vals) // 4, fp(vals), fp(out))
    assert (out >= vals).all()
    fin = np.isfinite(vals) & (vals > 1e-3) & (vals < 6e4)
    assert (out[fin] <= vals[fin] * (1 + 2.0 ** -10)).all()  # at most one FP16 ulp of slack
    assert np.isposinf(out[vals > 65504.0]).all() and np.isneginf(out[np.isneginf(vals)]).all()


def test_view_batch_chunking_host_logic():
    """msplat_b200/render.py: how a view batch is cut into chunks, and how chunks whose views hold more than
    2^31 - 1 tile intersections together are split again before anything is queued."""
    from msplat_b200.render import M_MAX, _chunks, _split_for_sort
    assert _chunks(8, 0) == [(0, 8)] and _chunks(8, 3) == [(0, 3), (3, 3), (6, 2)] and _chunks(2, 5) == [(0, 2)]
    assert M_MAX == 2 ** 31 - 1
    Ms = [10, 20, 30, 40]
    assert _split_for_sort([(0, 4)], Ms) == [(0, 4)]
    big = [2 ** 30, 2 ** 30, 5, 2 ** 30]           # 2^30 + 2^30 > 2^31 - 1: the first two views cannot share a sort
    assert _split_for_sort([(0, 4)], big) == [(0, 1), (1, 2), (3, 1)]
    assert _split_for_sort([(0, 2), (2, 2)], big) == [(0, 1), (1, 1), (2, 2)]
    import pytest
    with pytest.raises(RuntimeError, match="2\\^31"):
        _split_for_sort([(0, 1)], [2 ** 31])


def test_backward_blend_raw_moment_shift():
    """csrc/blend.cu::bwd_reduce_group accumulates, per pixel row of a warp's 8x4 block, the raw moments
    a = sum X, ax = sum X x, axx = sum X x^2 over the pixel columns x = 0..7 and shifts them once to the six moments
    about the Gaussian's centre (dx = ux - x, dy = uy - row).  The same algebra in float32 numpy against the direct
    float64 sums, for centres inside and far outside the block (the cancellation stays at rounding level because a
    far centre makes the moments themselves large)."""
    import numpy as np
    rng = np.random.default_rng(7)
    for ux, uy in ((3.3, 1.7), (-40.5, 12.25), (250.0, -180.0), (0.0, 0.0)):
        X = (rng.standard_normal((4, 8)) * rng.uniform(0.0, 1.0, (4, 8))).astype(np.float32)
        xs = np.arange(8, dtype=np.float32)
        m = np.zeros(6, dtype=np.float32)  # m0, mx, my, mxx, mxy, myy
        for h in range(2):  # the two half-warps own rows 2h and 2h + 1
            a = [np.float32(0)] * 2
            ax = [np.float32(0)] * 2
            axx = [np.float32(0)] * 2
            for r in range(2):
                for x in range(8):
                    v = X[2 * h + r, x]
                    a[r] = np.float32(a[r] + v)
                    ax[r] = np.float32(ax[r] + v * xs[x])
                    axx[r] = np.float32(axx[r] + v * xs[x] * xs[x])
            u = np.float32(ux)
            dy0 = np.float32(np.float32(uy) - np.float32(2 * h))
            dy1 = np.float32(dy0 - np.float32(1))
            s0, sx, sxx = np.float32(a[0] + a[1]), np.float32(ax[0] + ax[1]), np.float32(axx[0] + axx[1])
            part = np.array([s0, u * s0 - sx, dy0 * a[0] + dy1 * a[1], u * (u * s0 - 2 * sx) + sxx,
                             dy0 * (u * a[0] - ax[0]) + dy1 * (u * a[1] - ax[1]),
                             dy0 * dy0 * a[0] + dy1 * dy1 * a[1]], dtype=np.float32)
            m = (m + part).astype(np.float32)
        Xd = X.astype(np.float64)
        dx = ux - np.arange(8)[None, :]
        dy = uy - np.arange(4)[:, None]
        want = np.array([Xd.sum(), (Xd * dx).sum(), (Xd * dy).sum(), (Xd * dx * dx).sum(), (Xd * dx * dy).sum(),
                         (Xd * dy * dy).sum()])
        scale = np.array([np.abs(Xd).sum(), (np.abs(Xd) * np.abs(dx)).sum(), (np.abs(Xd) * np.abs(dy)).sum(),
                          (np.abs(Xd) * dx * dx).sum(), (np.abs(Xd) * np.abs(dx * dy)).sum(),
                          (np.abs(Xd) * dy * dy).sum()])
        assert np.all(np.abs(m - want) <= 2e-5 * scale + 1e-12), (ux, uy, m, want)
