"""BASELINE.json `configs` as GPU parity cases (the bench line is config #3; the others are checked
here against the unmodified reference CUDA build on identical inputs, at the named shapes or at a
size that runs in seconds, plus size-independent properties at the full size).

Bars: sort outputs / tiles_touched / radius bit-exact; images max abs 1e-4 x max(1, |image|);
gradients |d| <= 1e-3 |g| + K_NOISE x (the reference's measured run-to-run spread), see test_gpu_parity.
"""
import math

import pytest
import torch

from test_gpu_parity import DEV, K_NOISE, SH_ATOL, compare_grads, grad_close, spread
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
    target = torch.rand(3, sc.H, sc.W, generator=torch.Generator().manual_seed(2)).to(DEV)
    out = {}

    def run(api):
        L = mk()
        img = api.rasterization(*L, sc.intr, sc.extr, sc.W, sc.H, 1.0)
        out[api] = img.detach()
        torch.nn.functional.smooth_l1_loss(img, target).backward()      # the tutorial's loss (gs_2d.py:75)
        return [t.grad for t in L]

    # 64k-entry tile lists: every Gaussian's gradient is a sum over ~1e5 pixel pairs accumulated with float
    # atomics in the reference (alpha_blending.cu:218-243) -- the measured spread is what "equal" can mean
    compare_grads(lambda: run(ms), lambda: run(ref_msplat), ["dxyz", "dscale", "dquat", "dopacity", "drgb"],
                  "config2 init/ref")
    err = float((out[ms] - out[ref_msplat]).abs().max())
    assert err <= 1e-4, f"image max abs error vs reference {err}"


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
    _config4_compare(ms, ref_msplat, sc, "config4[30k,720p]")


def _config4_compare(ms, ref_msplat, sc, tag, fused=True):
    mk = lambda: [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, sc.shs)]
    C = int(sc.shs.shape[1])
    g = torch.randn(C, sc.H, sc.W, device=DEV)
    out = {}

    def run(key, fn):
        L = mk()
        img = fn(L)
        out[key] = img.detach()
        (img * g).sum().backward()
        grads = [t.grad for t in L]
        del L, img
        return grads

    names = ["dxyz", "dscale", "dquat", "dopacity", "dshs"]
    ref, nf = spread(lambda: run("ref", lambda L: steps_pipeline(ref_msplat, L, sc.intr, sc.extr, sc.W, sc.H, 0.0, False)), n=2)
    # dL_dshs and dL_dxyz flow through the degree-10 basis / its derivative: besides the atomics spread, the floor is
    # the reference's own compute_sh tolerance (SH_ATOL at the tensor's scale, see test_gpu_parity.SH_ATOL)
    nf = [max(f, SH_ATOL * float(r.abs().max())) if n in ("dxyz", "dshs") else f for n, f, r in zip(names, nf, ref)]
    ours = run("steps", lambda L: steps_pipeline(ms, L, sc.intr, sc.extr, sc.W, sc.H, 0.0, False))
    # 32 channels: dL_dalpha of a pair is a 32-term dot product, accumulated in packed pairs (FFMA2) and in two
    # 16-channel passes by our kernels, serially by the reference's -- a deterministic difference in summation order
    # that the run-to-run spread does not contain; measured need: 3.6-4.1 x the spread (dscale), hence 2 K_NOISE here
    # clamp_min(sh + 0.5, 0) is discontinuous: of the P x 32 colour values a handful lie within rounding of 0 and
    # pass the clamp in one library only (the degree-10 bases differ in evaluation order), which switches the 121
    # dL_dshs entries of that (Gaussian, channel) and the Gaussian's dL_dxyz on or off: all but 1e-5 of the elements
    frac = 1.0 - 1e-5
    for n, a, b, f in zip(names, ours, ref, nf):
        grad_close(a, b, noise=f, k=2 * K_NOISE, what=f"{tag} steps {n}", min_frac=frac)
    del ours
    if fused:
        ours = run("fused", lambda L: ms.rasterization_sh(*L, sc.intr, sc.extr, sc.W, sc.H, 0.0))
        for n, a, b, f in zip(names, ours, ref, nf):
            grad_close(a, b, noise=f, k=2 * K_NOISE, what=f"{tag} fused {n}", min_frac=frac)
        del ours
    assert out["steps"].shape == (C, sc.H, sc.W)
    scale = max(1.0, float(out["ref"].abs().max()))
    for name in ("steps", "fused") if fused else ("steps",):
        err = float((out[name] - out["ref"]).abs().max())
        assert err <= 1e-4 * scale, f"{name}: image max abs error vs reference {err}"


def test_config4_full_size_64bit_indexing_vs_reference(ms, ref_msplat):
    """BASELINE config #4 at a size whose SH tensor has more than 2^31 elements: 600k Gaussians x
    32 channels x 121 coefficients = 2.32e9 floats (9.3 GB) -- every SH index must be 64-bit.  Steps API of
    both libraries and our fused path on identical tensors, 1080p."""
    from msplat_b200.scenes import frustum_scene
    sc = frustum_scene(600_000, 1920, 1080, 2.0, seed=0, sh_degree=10, sh_channels=32).to(DEV)
    assert sc.shs.numel() > 2 ** 31
    _config4_compare(ms, ref_msplat, sc, "config4[600k,1080p,>2^31]")
    del sc
    torch.cuda.empty_cache()


def test_config5_4k_view_batch_vs_reference(ms, ref_msplat):
    """4K views (3840x2160, T = 32400 tiles, 47 key bits) of one cloud from the config-#5 cameras:
    the view-batch Function vs the reference's steps API view by view; sort outputs bit-exact."""
    from msplat_b200.scenes import frustum_scene, orbit_cameras
    sc = frustum_scene(400000, 3840, 2160, 3.0, seed=8, sh_degree=3).to(DEV)
    extrs = torch.stack([orbit_cameras(64)[k] for k in (0, 63)]).to(DEV)
    mk = lambda: [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, sc.shs)]
    A = mk()
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

    def run_ref():
        B = mk()
        for k in range(2):
            img_r = steps_pipeline(ref_msplat, B, sc.intr, extrs[k], sc.W, sc.H, 0.0, True)
            scale = max(1.0, float(img_r.detach().abs().max()))
            err = float((imgs[k].detach() - img_r.detach()).abs().max())
            assert err <= 1e-4 * scale, f"view {k}: image max abs error vs reference {err}"
            (img_r * g[k]).sum().backward()
        return [t.grad for t in B]

    ref, nf = spread(run_ref, n=1)
    for n, a, b, f in zip(["xyz", "scale", "quat", "opacity", "shs"], A, ref, nf):
        grad_close(a.grad, b, noise=f, what=f"config5 d{n}")


def test_config3_full_size_vs_reference(ms, ref_msplat):
    """The headline configuration at FULL size (3M Gaussians, 1920x1080, SH3, RGB+depth): our fused view
    batch (2 views, one launch per stage) against the unmodified reference build's steps API on identical
    tensors -- tiles_touched / radius / idx_sorted / tile_range bit-exact per view, image max abs 1e-4
    (x the depth channel's scale), gradients within 1e-3 |g| + K_NOISE x the reference's own spread."""
    from msplat_b200.scenes import frustum_scene, orbit_cameras
    sc = frustum_scene(3_000_000, 1920, 1080, 2.0, seed=0, sh_degree=3).to(DEV)
    extrs = torch.stack([e for e in orbit_cameras(8, yaw_deg=20.0, shift=1.0)[:2]]).to(DEV)
    mk = lambda: [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, sc.shs)]
    g = torch.randn(2, 4, sc.H, sc.W, device=DEV)
    import copy
    for k in range(2):
        sk = copy.copy(sc)
        sk.extr = extrs[k]
        ours, ref = _geometry(ms, sk), _geometry(ref_msplat, sk)
        for name, a, b in zip(["uv", "depth", "conic", "radius", "tiles", "idx_sorted", "tile_range"], ours, ref):
            assert torch.equal(a, b), f"view {k}: {name} differs from the reference"
        # the batched sort of the fused path: same per-view order, ids offset by view * P', ranges by the batch
        del ours, ref
    A = mk()
    imgs = ms.rasterization_sh_views(*A, sc.intr, extrs, sc.W, sc.H, 0.0, with_depth=True)
    (imgs * g).sum().backward()

    def run_ref():
        B = mk()
        for k in range(2):
            img_r = steps_pipeline(ref_msplat, B, sc.intr, extrs[k], sc.W, sc.H, 0.0, True)
            scale = max(1.0, float(img_r.detach().abs().max()))
            err = float((imgs[k].detach() - img_r.detach()).abs().max())
            assert err <= 1e-4 * scale, f"view {k}: image max abs error vs reference {err} (scale {scale})"
            (img_r * g[k]).sum().backward()
        return [t.grad for t in B]

    ref, nf = spread(run_ref, n=1)
    for n, a, b, f in zip(["xyz", "scale", "quat", "opacity", "shs"], A, ref, nf):
        grad_close(a.grad, b, noise=f, what=f"config3 full size d{n}")


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

    def backward(cot):
        for p in P:
            p.grad = None
        (img * cot).sum().backward(retain_graph=True)
        return [p.grad.clone() for p in P]

    # three backward passes through the same graph: two with g1 (their difference is the run-to-run spread of
    # our atomics-based backward blend = the noise floor), one with 2 g1 (linearity)
    ga, gb = backward(g1), backward(g1)
    g2 = backward(2.0 * g1)
    for n, a, b, c in zip(["xyz", "scale", "quat", "opacity", "shs"], ga, gb, g2):
        assert bool(torch.isfinite(c).all())
        grad_close(c, 2.0 * a, noise=2.0 * float((a - b).abs().max()), what=f"config3 linearity d{n}")
