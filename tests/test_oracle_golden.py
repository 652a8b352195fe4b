"""The oracle pinned against the reference's own golden vectors / restatements (CPU only).

Fixtures come from tests/golden/make_golden.py, which runs the reference's torch
restatements (test/test_*.py) and parses the reference CUDA text for the SH polynomials.
"""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN

T = torch.from_numpy


def test_sort_known_answer():
    """/root/reference/test/test_sort_gaussian.py:9-52"""
    ka = json.load(open(os.path.join(GOLDEN, "sort_known_answer.json")))
    uv = torch.tensor(ka["uv"], dtype=torch.float32)
    depth = torch.tensor(ka["depth"], dtype=torch.float32)[:, None]
    radius = torch.tensor(ka["radius"], dtype=torch.int32)[:, None]
    tiles = torch.tensor(ka["tiles"], dtype=torch.int32)[:, None]
    ids, tr = oracle.sort_gaussian(uv, depth, ka["W"], ka["H"], radius, tiles)
    assert ids.tolist() == ka["idx_sorted"]
    assert tr.tolist() == ka["tile_range"]


def test_sort_phantom_and_empty():
    # tiles > 0 with radius == 0 leaves zero-initialised (key 0, idx 0) slots: SURVEY H3
    uv = torch.tensor([[20.0, 4.0], [5.0, 5.0]])
    depth = torch.tensor([[1.0], [2.0]])
    radius = torch.tensor([0, 2], dtype=torch.int32)
    tiles = torch.tensor([2, 1], dtype=torch.int32)
    ids, tr = oracle.sort_gaussian(uv, depth, 32, 16, radius, tiles)
    assert ids.tolist() == [0, 0, 1]
    assert tr.tolist() == [[0, 3], [0, 0]]
    ids, tr = oracle.sort_gaussian(torch.zeros(0, 2), torch.zeros(0, 1), 32, 16, torch.zeros(0, dtype=torch.int32),
                                   torch.zeros(0, dtype=torch.int32))
    assert ids.numel() == 0 and tr.tolist() == [[0, 0], [0, 0]]


def test_sh_basis_matches_reference_text_and_torch(golden):
    g = golden("sh_basis.npz")
    dirs = T(g["dirs"])
    B = oracle.sh_basis(dirs, 121).numpy()
    for key in ("basis_cuda_text", "basis_torch"):
        ref = g[key]
        scale = np.maximum(np.abs(ref), 1.0)
        assert np.max(np.abs(B - ref) / scale) < 1e-12, key


def test_sh_basis_matches_scipy_on_sphere():
    """Same cross-check as /root/reference/test/test_compute_sh.py:328-376 (scipy renamed
    sph_harm -> sph_harm_y)."""
    sp = pytest.importorskip("scipy.special")
    rng = np.random.default_rng(1)
    d = rng.normal(size=(16, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    polar = np.arccos(d[:, 2])
    azim = np.arctan2(d[:, 1], d[:, 0])
    B = oracle.sh_basis(T(d), 121).numpy()
    for n in range(11):
        for m in range(-n, n + 1):
            if hasattr(sp, "sph_harm_y"):
                y1, y2 = sp.sph_harm_y(n, m, polar, azim), sp.sph_harm_y(n, -m, polar, azim)
            else:
                y1, y2 = sp.sph_harm(m, n, azim, polar), sp.sph_harm(-m, n, azim, polar)
            if m < 0:
                val = ((-1.0) ** (-m)) * (1j * np.sqrt(0.5) * (y1 - ((-1.0) ** (-m)) * y2)).real
            elif m > 0:
                val = ((-1.0) ** m) * (np.sqrt(0.5) * (y2 + ((-1.0) ** m) * y1)).real
            else:
                val = y1.real
            assert np.allclose(B[:, n * n + n + m], val, atol=1e-10), (n, m)


def test_compute_sh_golden(golden):
    g = golden("compute_sh.npz")
    dirs = T(g["dirs"]).requires_grad_()
    shs = T(g["shs"]).requires_grad_()
    val = oracle.compute_sh(shs, dirs)
    val.mean().backward()
    assert np.allclose(val.detach().numpy(), g["value"], rtol=1e-10, atol=1e-10)
    assert np.allclose(shs.grad.numpy(), g["dshs"], rtol=1e-10, atol=1e-12)
    assert np.allclose(dirs.grad.numpy(), g["ddirs"], rtol=1e-9, atol=1e-10)


def test_project_point_golden(golden):
    g = golden("project_point.npz")
    xyz = T(g["xyz"]).requires_grad_()
    uv, depth = oracle.project_point(xyz, T(g["intr"]), T(g["extr"]), int(g["W"]), int(g["H"]),
                                     float(g["nearest"]), float(g["extent"]))
    torch.testing.assert_close(uv.detach(), T(g["uv"]), rtol=1e-5, atol=2e-3)  # pixels ~1e3, f32 matmul order
    torch.testing.assert_close(depth.detach(), T(g["depth"]), rtol=1e-5, atol=1e-4)
    assert ((depth == 0) == (T(g["depth"]) == 0)).all()
    ((uv * T(g["guv"])).sum() + (depth * T(g["gd"])).sum()).backward()
    torch.testing.assert_close(xyz.grad, T(g["dxyz"]), rtol=1e-4, atol=1e-4)


def test_compute_cov3d_golden(golden):
    g = golden("compute_cov3d.npz")
    s, q = T(g["scale"]).requires_grad_(), T(g["quat"]).requires_grad_()
    cov = oracle.compute_cov3d(s, q)
    torch.testing.assert_close(cov.detach(), T(g["cov3d"]))
    (cov * T(g["g"])).sum().backward()
    torch.testing.assert_close(s.grad, T(g["dscale"]), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(q.grad, T(g["dquat"]), rtol=1e-5, atol=1e-5)


def test_ewa_project_golden(golden):
    g = golden("ewa_project.npz")
    xyz, cov = T(g["xyz"]).requires_grad_(), T(g["cov3d"]).requires_grad_()
    intr, extr = T(g["intr"]).requires_grad_(), T(g["extr"]).requires_grad_()
    conic, radius, tiles = oracle.ewa_project(xyz, cov, intr, extr, T(g["uv"]), int(g["W"]), int(g["H"]),
                                              T(g["visible"]))
    assert (radius.numpy() == g["radius"]).all()   # exact, as test/test_ewa_project.py:230-231
    assert (tiles.numpy() == g["tiles"]).all()
    torch.testing.assert_close(conic.detach(), T(g["conic"]), rtol=1e-4, atol=1e-6)
    conic.sum().backward()
    torch.testing.assert_close(xyz.grad, T(g["dxyz"]), rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(cov.grad, T(g["dcov3d"]), rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(intr.grad, T(g["dintr"]), rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(extr.grad, T(g["dextr"]), rtol=1e-3, atol=1e-5)


def test_alpha_blending_golden(golden):
    """Reference loop restatement test/test_alpha_blending.py:6-63 (run by make_golden.py)."""
    g = golden("alpha_blending.npz")
    uv, conic = T(g["uv"]).requires_grad_(), T(g["conic"]).requires_grad_()
    op, feat = T(g["opacity"]).requires_grad_(), T(g["feature"]).requires_grad_()
    W, H, bg = int(g["W"]), int(g["H"]), float(g["bg"])
    ids, tr = oracle.sort_gaussian(uv, T(g["depth"]), W, H, T(g["radius"]), T(g["tiles"]))
    assert (ids.numpy() == g["idx_sorted"]).all() and (tr.numpy() == g["tile_range"]).all()
    img = oracle.alpha_blending(uv, conic, op, feat, ids, tr, bg, W, H)
    torch.testing.assert_close(img.detach(), T(g["image"]), rtol=1e-5, atol=1e-5)
    img.sum().backward()
    torch.testing.assert_close(uv.grad, T(g["duv"]), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(conic.grad, T(g["dconic"]), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(op.grad, T(g["dopacity"]), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(feat.grad, T(g["dfeature"]), rtol=1e-3, atol=1e-4)


def test_blend_c_matches_python_loop():
    torch.manual_seed(3)
    N, C, W, H = 12, 3, 32, 16
    uv = torch.rand(N, 2) * torch.tensor([W, H])
    A = torch.randn(N, 2, 2) * 0.3
    cv = A @ A.transpose(1, 2) + 0.02 * torch.eye(2)
    conic = torch.stack([cv[:, 0, 0], cv[:, 0, 1], cv[:, 1, 1]], -1)
    op, feat = torch.rand(N, 1), torch.rand(N, C)
    depth = torch.rand(N, 1) * 5 + 0.1
    radius = torch.full((N,), 40, dtype=torch.int32)
    xmin, ymin, xmax, ymax = oracle.get_rect(uv, radius, W, H)
    tiles = ((xmax - xmin) * (ymax - ymin)).int()
    ids, tr = oracle.sort_gaussian(uv, depth, W, H, radius, tiles)
    a = oracle.alpha_blending(uv, conic, op, feat, ids, tr, 0.5, W, H)
    b = oracle.alpha_blending_loop(uv, conic, op, feat, ids, tr, 0.5, W, H)
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)


def test_rasterization_oracle_runs():
    from msplat_b200.scenes import cube_scene
    sc = cube_scene(P=500, W=64, H=48)
    feat = torch.rand(500, 3)
    img = oracle.rasterization(sc.xyz, sc.scale, sc.quat, sc.opacity, feat, sc.intr, sc.extr, sc.W, sc.H, 0.0)
    assert img.shape == (3, 48, 64) and torch.isfinite(img).all() and img.abs().sum() > 0
