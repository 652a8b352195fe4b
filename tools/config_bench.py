#!/usr/bin/env python3
"""Full-size timings and parity of the BASELINE configs that are not the bench line (1x B200):

  #2  gs_2d initialisation, 100k Gaussians, 512x512 RGB: rasterization() fwd+bwd (sort-dominated: M ~ 65M keys)
  #4  1M Gaussians, 1080p, SH degree 10, 32 channels (15.5 GB of coefficients): fwd+bwd
  #5  6M Gaussians, one 3840x2160 view, SH degree 3, RGB+depth: fwd+bwd

Each with the unmodified reference build (steps API) and with msplat_b200 (steps API and fused
path); images are compared at full size (max abs error, the 1e-4 bar of the parity tests) and the
sort outputs bit for bit.  Prints one JSON object per config.  `python tools/config_bench.py [2 4 5]`.
"""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import msplat_b200 as ms  # noqa: E402
from msplat_b200.scenes import bunny2d_scene, frustum_scene, orbit_cameras  # noqa: E402

DEV = "cuda:0"


def reference():
    p = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(p, "msplat")):
        return None
    sys.path.insert(0, p)
    import msplat
    return msplat


def timed(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def steps(api, P, intr, extr, W, H, with_depth, G):
    xyz, scale, quat, opacity, shs = P
    uv, depth = api.project_point(xyz, intr, extr, W, H)
    vis = depth != 0
    R, t = extr[:3, :3], extr[:3, 3]
    dirs = xyz - (-(R.T @ t))
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    col = torch.clamp_min(api.compute_sh(shs, dirs, vis.squeeze(-1)) + 0.5, 0.0)
    feat = torch.cat([col, depth], dim=-1) if with_depth else col
    cov = api.compute_cov3d(scale, quat, vis)
    conic, radius, tiles = api.ewa_project(xyz, cov, intr, extr, uv, W, H, vis)
    ids, tr = api.sort_gaussian(uv, depth, W, H, radius, tiles)
    img = api.alpha_blending(uv, conic, opacity, feat, ids, tr, 0.0, W, H)
    if G is not None:
        (img * G).sum().backward()
    return img, ids, tr


def sh_config(name, Pn, W, H, sigma, deg, Cs, with_depth, extr=None):
    ref = reference()
    sc = frustum_scene(Pn, W, H, sigma, seed=0, sh_degree=0, with_sh=False).to(DEV)
    D = (deg + 1) ** 2
    g = torch.Generator(device=DEV).manual_seed(5)
    shs = 0.1 * torch.randn(Pn, Cs, D, device=DEV, generator=g)
    shs[:, :, 0] *= 5.0
    extr = sc.extr if extr is None else extr.to(DEV)
    C = Cs + (1 if with_depth else 0)
    G = torch.randn(C, H, W, device=DEV, generator=g)
    P = [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, shs)]
    del shs

    def zero():
        for p in P:
            p.grad = None

    out = {"config": name, "gaussians": Pn, "size": [W, H], "sh_degree": deg, "channels": C}
    with torch.no_grad():
        img_o, ids_o, tr_o = steps(ms, P, sc.intr, extr, W, H, with_depth, None)
        out["keys"] = int(ids_o.numel())
        if ref is not None:
            img_r, ids_r, tr_r = steps(ref, P, sc.intr, extr, W, H, with_depth, None)
            out["sort_bit_exact"] = bool(torch.equal(ids_o, ids_r) and torch.equal(tr_o, tr_r))
            out["image_max_abs_err_steps"] = float((img_o - img_r).abs().max())
            img_f = ms.rasterization_sh(*P, sc.intr, extr, W, H, 0.0, with_depth=with_depth)
            out["image_max_abs_err_fused"] = float((img_f - img_r).abs().max())
            out["image_scale"] = float(img_r.abs().max())
            del img_r, ids_r, tr_r, img_f
        del img_o, ids_o, tr_o
    torch.cuda.empty_cache()

    def run_steps(api):
        zero()
        steps(api, P, sc.intr, extr, W, H, with_depth, G)

    def run_fused():
        zero()
        img = ms.rasterization_sh(*P, sc.intr, extr, W, H, 0.0, with_depth=with_depth)
        (img * G).sum().backward()

    from msplat_b200 import _lib
    run_fused()
    torch.cuda.synchronize()
    _lib.TIMING = []
    run_fused()
    torch.cuda.synchronize()
    tl, _lib.TIMING = _lib.TIMING, None
    stages = {}
    for nm, a, b in tl:
        stages[nm] = round(stages.get(nm, 0.0) + a.elapsed_time(b), 4)
    out["fused_stage_ms"] = stages
    out["ms_fwd_bwd"] = {"msplat_b200 fused": round(timed(run_fused), 3),
                         "msplat_b200 steps API": round(timed(lambda: run_steps(ms)), 3)}
    if ref is not None:
        out["ms_fwd_bwd"]["reference build steps API"] = round(timed(lambda: run_steps(ref)), 3)
    print(json.dumps(out), flush=True)


def config2():
    ref = reference()
    sc = bunny2d_scene(100000, 512, 512, seed=123).to(DEV)
    rgb = torch.sigmoid(torch.rand(100000, 3, generator=torch.Generator().manual_seed(1))).to(DEV)
    target = torch.rand(3, 512, 512, generator=torch.Generator().manual_seed(2)).to(DEV)
    P = [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, rgb)]

    def run(api):
        for p in P:
            p.grad = None
        img = api.rasterization(*P, sc.intr, sc.extr, 512, 512, 1.0)
        torch.nn.functional.smooth_l1_loss(img, target).backward()
        return img

    out = {"config": "#2 gs_2d initialisation", "gaussians": 100000, "size": [512, 512], "channels": 3}
    with torch.no_grad():
        uv, depth = ms.project_point(sc.xyz, sc.intr, sc.extr, 512, 512)
        cov = ms.compute_cov3d(sc.scale, sc.quat, depth != 0)
        _, radius, tiles = ms.ewa_project(sc.xyz, cov, sc.intr, sc.extr, uv, 512, 512, depth != 0)
        out["keys"] = int(tiles.sum())
    out["ms_fwd_bwd"] = {"msplat_b200 rasterization()": round(timed(lambda: run(ms)), 3)}
    # per-C-ABI-call durations of one fwd+bwd (CUDA events on the launching stream)
    from msplat_b200 import _lib
    _lib.TIMING = []
    run(ms)
    torch.cuda.synchronize()
    tl, _lib.TIMING = _lib.TIMING, None
    stages = {}
    for name, a, b in tl:
        stages[name] = round(stages.get(name, 0.0) + a.elapsed_time(b), 4)
    out["stage_ms"] = stages
    if "sort_gaussian" in stages:
        out["sort_Gkeys_per_s"] = round(out["keys"] / (stages["sort_gaussian"] * 1e-3) / 1e9, 1)
    if ref is not None:
        out["image_max_abs_err"] = float((run(ms).detach() - run(ref).detach()).abs().max())
        out["ms_fwd_bwd"]["reference build rasterization()"] = round(timed(lambda: run(ref)), 3)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["2", "4", "5"]
    if "2" in which:
        config2()
    if "4" in which:
        sh_config("#4 high-order SH + wide features", 1_000_000, 1920, 1080, 2.0, 10, 32, False)
    if "5" in which:
        sh_config("#5 one 4K view of the 6M cloud", 6_000_000, 3840, 2160, 3.0, 3, 3, True, extr=orbit_cameras(64)[17])
