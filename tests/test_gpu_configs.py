"""BASELINE.json `configs` as GPU parity cases (the bench line is config #3; the others are checked
here against the unmodified reference CUDA build on identical inputs, at the named shapes or at a
size that runs in seconds, plus size-independent properties at the full size).

Bars: sort outputs / tiles_touched / radius bit-exact; images max abs 1e-4 x max(1, |image|);
gradients |d| <= rel |g| + eps max|g|.
"""
import math

import pytest
import torch

from test_gpu_parity import DEV, grad_close
from test_gpu_render_sh import steps_pipeline

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ms():
    import msplat_b200
    return msplat_b200


def _geometry(api, sc):
    uv, depth = api.project_point(sc.xyz, sc.intr, sc.extr, sc.W, sc.H)
    vis = depth != 0
    cov = api.compute_cov3d(sc.scale, sc.quat, vis)
    conic, radius, tiles = api.ewa_project(sc.xyz, cov, sc.intr, sc.extr, uv, sc.W, sc.H, vis)
    ids, tr = api.sort_gaussian(uv, depth, sc.W, sc.H, radius, tiles)
    return uv, depth, conic, radius, tiles, ids, tr


def test_config2_bunny2d_init_vs_reference(ms, ref_msplat):
    """tutorials/gs_2d.py at initialisation with num_points = 100000 (BASELINE config #2): a
    sort-dominated case -- ~650 tiles per Gaussian, M ~ 65M keys, 64k-entry tile lists."""
    from msplat_b200.scenes import bunny2d_scene
    sc = bunny2d_scene(100000, 512, 512, seed=123).to(DEV)
    ours, ref = _geometry(ms, sc), _geometry(ref_msplat, sc)
    for name, a, b in zip(["uv", "depth", "conic", "radius", "tiles", "idx_sorted", "tile_range"], ours, ref):
        assert torch.equal(a, b), f"{name} differs from the reference"
    assert ours[5].numel() > 30_000_000
    rgb = torch.sigmoid(torch.rand(sc.xyz.shape[0], 3, generator=torch.Generator().manual_seed(1))).to(DEV)
    mk = lambda: [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, rgb)]
    A, B = mk(), mk()
    img = ms.rasterization(*A, sc.intr, sc.extr, sc.W, sc.H, 1.0)
    img_r = ref_msplat.rasterization(*B, sc.intr, sc.extr, sc.W, sc.H, 1.0)
    err = float((img.detach() - img_r.detach()).abs().max())
    assert err <= 1e-4, f"image max abs error vs reference {err}"
    target = torch.rand(3, sc.H, sc.W, generator=torch.Generator().manual_seed(2)).to(DEV)
    torch.nn.functional.smooth_l1_loss(img, target).backward()      # the tutorial's loss (gs_2d.py:75)
    torch.nn.functional.smooth_l1_loss(img_r, target).backward()
    for n, a, b in zip(["xyz", "scale", "quat", "opacity", "rgb"], A, B):
        grad_close(a.grad, b.grad, rel=5e-3, eps=1e-3, what=f"config2 d{n}")


def test_config2_bunny2d_training_steps_track_reference(ms, ref_msplat):
    """A few Adam steps of the gs_2d.py loop (lr 1e-2... as in the tutorial) on 20k points: the loss
    curves of the two libraries stay together (drop-in at the application level)."""
    from msplat_b200.scenes import bunny2d_scene
    sc = bunny2d_scene(20000, 256, 256, seed=123).to(DEV)
    target = torch.rand(3, sc.H, sc.W, generator=torch.Generator().manual_seed(5)).to(DEV)
    g = torch.Generator().manual_seed(7)
    rgb0 = torch.rand(sc.xyz.shape[0], 3, generator=g).to(DEV)

    def run(api):
        P = [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, rgb0)]
        opt = torch.optim.Adam(P, lr=1e-3)
        losses = []
        for _ in range(8):
            xyz, scale, quat, op, rgb = P
            img = api.rasterization(xyz, scale.abs() + 1e-8, quat / quat.norm(dim=-1, keepdim=True), torch.sigmoid(op),
                                    torch.sigmoid(rgb), sc.intr, sc.extr, sc.W, sc.H, 1.0)
            loss = torch.nn.functional.smooth_l1_loss(img, target)
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss))
        return losses

    a, b = run(ms), run(ref_msplat)
    assert a[-1] < a[0], "the loss must go down"
    for x, y in zip(a, b):
        assert abs(x - y) <= 2e-3 * abs(y) + 1e-6, f"loss curves diverge: {a} vs {b}"


def test_config4_high_order_sh_wide_features_vs_reference(ms, ref_msplat):
    """SH degree 10 + 32-channel feature maps (BASELINE config #4) at 1280x720, 30k Gaussians:
    steps API of both libraries, and our fused path on the same scene."""
    from msplat_b200.scenes import frustum_scene
    sc = frustum_scene(30000, 1280, 720, 3.0, seed=6, sh_degree=10, sh_channels=32).to(DEV)
    mk = lambda: [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, sc.shs)]
    A, B, F = mk(), mk(), mk()
    img = steps_pipeline(ms, A, sc.intr, sc.extr, sc.W, sc.H, 0.0, False)
    img_r = steps_pipeline(ref_msplat, B, sc.intr, sc.extr, sc.W, sc.H, 0.0, False)
    img_f = ms.rasterization_sh(*F, sc.intr, sc.extr, sc.W, sc.H, 0.0)
    assert img.shape == (32, sc.H, sc.W)
    scale = max(1.0, float(img_r.detach().abs().max()))
    for name, x in (("steps", img), ("fused", img_f)):
        err = float((x.detach() - img_r.detach()).abs().max())
        assert err <= 1e-4 * scale, f"{name}: image max abs error vs reference {err}"
    g = torch.randn(32, sc.H, sc.W, device=DEV)
    for x in (img, img_r, img_f):
        (x * g).sum().backward()
    for n, a, b, f in zip(["xyz", "scale", "quat", "opacity", "shs"], A, B, F):
        grad_close(a.grad, b.grad, rel=5e-3, eps=5e-4, what=f"config4 steps d{n}")
        grad_close(f.grad, b.grad, rel=5e-3, eps=5e-4, what=f"config4 fused d{n}")


def test_config5_4k_view_batch_vs_reference(ms, ref_msplat):
    """4K views (3840x2160, T = 32400 tiles, 47 key bits) of one cloud from the config-#5 cameras:
    the view-batch Function vs the reference's steps API view by view; sort outputs bit-exact."""
    from msplat_b200.scenes import frustum_scene, orbit_cameras
    sc = frustum_scene(400000, 3840, 2160, 3.0, seed=8, sh_degree=3).to(DEV)
    extrs = torch.stack([orbit_cameras(64)[k] for k in (0, 63)]).to(DEV)
    mk = lambda: [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, sc.shs)]
    A, B = mk(), mk()
    imgs = ms.rasterization_sh_views(*A, sc.intr, extrs, sc.W, sc.H, 0.0, with_depth=True)
    g = torch.randn(2, 4, sc.H, sc.W, device=DEV)
    (imgs * g).sum().backward()
    for k in range(2):
        import copy
        sk = copy.copy(sc)
        sk.extr = extrs[k]
        ours, ref = _geometry(ms, sk), _geometry(ref_msplat, sk)
        for name, a, b in zip(["uv", "depth", "conic", "radius", "tiles", "idx_sorted", "tile_range"], ours, ref):
            assert torch.equal(a, b), f"view {k}: {name} differs from the reference"
        img_r = steps_pipeline(ref_msplat, B, sc.intr, extrs[k], sc.W, sc.H, 0.0, True)
        scale = max(1.0, float(img_r.detach().abs().max()))
        err = float((imgs[k].detach() - img_r.detach()).abs().max())
        assert err <= 1e-4 * scale, f"view {k}: image max abs error vs reference {err}"
        (img_r * g[k]).sum().backward()
    for n, a, b in zip(["xyz", "scale", "quat", "opacity", "shs"], A, B):
        grad_close(a.grad, b.grad, rel=5e-3, eps=5e-4, what=f"config5 d{n}")


def test_config3_full_size_properties(ms):
    """BASELINE config #3 at full size (3M Gaussians, 1080p, SH3, RGB+depth) through the fused path:
    properties that do not need a second implementation -- finite image, alpha-compositing bounds
    (every colour is a convex combination of clamped colours, so 0 <= rgb), linearity of the
    backward pass in the cotangent, and determinism of the forward pass."""
    from msplat_b200.scenes import frustum_scene
    sc = frustum_scene(3_000_000, 1920, 1080, 2.0, seed=0, sh_degree=3).to(DEV)
    P = [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, sc.shs)]
    img = ms.rasterization_sh(*P, sc.intr, sc.extr, sc.W, sc.H, 0.0, with_depth=True)
    assert bool(torch.isfinite(img).all()) and float(img[:3].min()) >= 0.0
    assert float(img[3].max()) <= 50.0 * 1.001  # depth channel: convex combination of z in [1, 50]
    img2 = ms.rasterization_sh(*[p.detach() for p in P], sc.intr, sc.extr, sc.W, sc.H, 0.0, with_depth=True)
    assert torch.equal(img, img2)
    g1 = torch.randn(4, sc.H, sc.W, device=DEV)
    (img * g1).sum().backward(retain_graph=True)
    ga = [p.grad.clone() for p in P]
    for p in P:
        p.grad = None
    (img * (2.0 * g1)).sum().backward()
    for n, a, p in zip(["xyz", "scale", "quat", "opacity", "shs"], ga, P):
        assert bool(torch.isfinite(p.grad).all())
        grad_close(p.grad, 2.0 * a, rel=1e-3, eps=1e-4, what=f"linearity d{n}")
