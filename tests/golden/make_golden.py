#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the reference itself.

Run in the BUILD CONTAINER only (needs /root/reference and, for the torch restatements, the
reference package importable from baseline/_ref); the GPU box never runs this, it only reads
the committed fixtures.

    python tests/golden/make_golden.py

Fixtures written:
  sort_known_answer.json   the reference's own known-answer test
                           (/root/reference/test/test_sort_gaussian.py:9-52)
  sh_basis.npz             SH basis values for degree 0..10 at on- and OFF-sphere directions from
                           (a) the polynomial text of msplat/src/compute_sh.cu:116-503 evaluated in
                           float64 and (b) the reference's torch restatement
                           test/test_compute_sh.py:162-325 (eval_sh_bases)
  project_point.npz, compute_cov3d.npz, ewa_project.npz, compute_sh.npz, alpha_blending.npz
                           inputs / outputs / autograd gradients of the reference's torch
                           restatements (test/test_*.py) run on CPU (``.cuda()`` patched to a no-op)
                           at reduced N so the files stay small
  gs2d_target.png          the 512x512 RGB target of the reference's 2-D fitting tutorial
                           (tutorials/gs_2d.py:51 reads data/stanford-bunny.jpg), decoded once and stored
                           losslessly so that the acceptance run (BASELINE config #2) fits the same pixels
"""
import json
import math
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def parse_cuda_sh():
    """Turn evaluateSH{l}Forward (compute_sh.cu:116-503) into python callables basis(x,y,z)->[121]."""
    src = open(os.path.join(REF, "msplat/src/compute_sh.cu")).read()
    # constants
    consts = {}
    m = re.search(r"SH_C0 = ([0-9.eE+-]+)f;", src)
    consts["SH_C0"] = float(m.group(1))
    for m in re.finditer(r"SH_C(\d+)\[\] = \{([^}]*)\}", src):
        vals = [float(v.strip().rstrip("f")) for v in m.group(2).split(",") if v.strip()]
        consts["SH_C%s" % m.group(1)] = vals
    terms = {}  # index -> (python expr, local-defs)
    for l in range(0, 11):
        m = re.search(r"float evaluateSH%dForward\((.*?)\n}\n" % l, src, re.S)
        body = m.group(0)
        body = re.sub(r"(\d)f\b", r"\1", body)  # 3.0f -> 3.0
        stmts = [re.sub(r"\s+", " ", s.strip()) for s in body.split(";")]
        defs = []
        for s in stmts:
            if s.startswith("float ") and "result" not in s and "evaluateSH" not in s:
                for d in s[len("float "):].split(","):
                    name, expr = d.split("=")
                    defs.append((name.strip(), expr.strip()))
            mm = re.match(r"result \+= sh\[(\d+)\] \* (.*)$", s)
            if mm:
                terms[int(mm.group(1))] = (mm.group(2), list(defs))
            mm = re.search(r"return SH_C0 \* sh\[0\]", s)
            if mm:
                terms[0] = ("SH_C0", [])
    assert sorted(terms) == list(range(121)), sorted(terms)

    def basis(x, y, z):
        out = np.zeros(x.shape + (121,), dtype=np.float64)
        for i, (expr, defs) in terms.items():
            env = {"x": x, "y": y, "z": z}
            env.update(consts)
            env["dir"] = None
            for name, e in defs:
                if name in ("x", "y", "z"):
                    continue
                env[name] = eval(e, {}, env)
            out[..., i] = eval(expr, {}, env)
        return out

    return basis


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)

    # ---------------- sort known answer ----------------
    ka = {
        "W": 32, "H": 16,
        "uv": [[2, 2], [30, 2], [8, 8], [30, 2]],
        "depth": [1.0, 2.0, 1.5, 3.0],
        "radius": [2, 8, 16, 1],
        "tiles": [1, 1, 2, 1],
        "idx_sorted": [0, 2, 2, 1, 3],
        "tile_range": [[0, 2], [2, 5]],
        "source": "/root/reference/test/test_sort_gaussian.py:9-52",
    }
    json.dump(ka, open(os.path.join(HERE, "sort_known_answer.json"), "w"), indent=1)

    # ---------------- SH basis from the CUDA text ----------------
    basis = parse_cuda_sh()
    dirs_on = rng.normal(size=(64, 3))
    dirs_on /= np.linalg.norm(dirs_on, axis=-1, keepdims=True)
    dirs_off = rng.normal(size=(64, 3)) * 0.8
    dirs = np.concatenate([dirs_on, dirs_off], 0)
    B_text = basis(dirs[:, 0], dirs[:, 1], dirs[:, 2])

    # ---------------- reference torch restatements ----------------
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    sys.path.insert(0, os.path.join(REF, "test"))
    torch.Tensor.cuda = lambda self, *a, **k: self  # the restatements hard-code .cuda()
    import scipy.special as _sp
    if not hasattr(_sp, "sph_harm"):  # scipy >= 1.17 removed it; the reference test imports it at module level
        _sp.sph_harm = lambda m, n, theta, phi: _sp.sph_harm_y(n, m, phi, theta)
    import test_compute_sh as tsh

    B_torch = tsh.eval_sh_bases(121, torch.from_numpy(dirs)).numpy()
    np.savez_compressed(os.path.join(HERE, "sh_basis.npz"), dirs=dirs, basis_cuda_text=B_text, basis_torch=B_torch)
    on = slice(0, 64)
    print("SH: |text - torch| on sphere  max", np.abs(B_text[on] - B_torch[on]).max())
    print("SH: |text - torch| off sphere max", np.abs(B_text[64:] - B_torch[64:]).max())

    # project_point (test/test_project_points.py:8-52, seed 123, DTU-like camera :67-72)
    import test_project_points as tpp
    torch.manual_seed(123)
    N, W, H = 300, 1600, 1200
    intr = torch.tensor([2892.33, 2883.18, 823.205, 619.071])
    extr = torch.tensor([[0.970263, 0.00747983, 0.241939, -191.02],
                         [-0.0147429, 0.999493, 0.0282234, 3.2883],
                         [-0.241605, -0.030951, 0.969881, 22.5401]])
    xyz = (torch.rand(N, 3) * 2 - 1) * 400
    xyz[:, 2] = xyz[:, 2].abs() * 2 + 100
    x1 = xyz.clone().requires_grad_()
    uv, depth = tpp.project_point_torch_impl(x1, intr, extr, W, H, nearest=0.2, extent=1.3)
    guv, gd = torch.randn(N, 2), torch.randn(N, 1)
    ((uv * guv).sum() + (depth * gd).sum()).backward()
    np.savez_compressed(os.path.join(HERE, "project_point.npz"), xyz=xyz.numpy(), intr=intr.numpy(), extr=extr.numpy(),
                        W=W, H=H, nearest=0.2, extent=1.3, uv=uv.detach().numpy(), depth=depth.detach().numpy(),
                        guv=guv.numpy(), gd=gd.numpy(), dxyz=x1.grad.numpy())

    # compute_cov3d (test/test_compute_cov3d.py:7-36)
    import test_compute_cov3d as tcc
    N = 300
    s = (torch.rand(N, 3) + 0.1).requires_grad_()
    q = torch.rand(N, 4)
    q = (q / q.norm(dim=-1, keepdim=True)).requires_grad_()
    cov = tcc.compute_cov3d_torch_impl(s, q)
    g = torch.randn(N, 6)
    (cov * g).sum().backward()
    np.savez_compressed(os.path.join(HERE, "compute_cov3d.npz"), scale=s.detach().numpy(), quat=q.detach().numpy(),
                        cov3d=cov.detach().numpy(), g=g.numpy(), dscale=s.grad.numpy(), dquat=q.grad.numpy())

    # ewa_project (test/test_ewa_project.py:11-107, camera :148-154, seed 123)
    import test_ewa_project as tep
    torch.manual_seed(123)
    N, W, H = 400, 800, 800
    intr = torch.tensor([1111.0, 1111.0, H / 2, W / 2 / 2])
    extr = torch.tensor([[6.1182e-01, 7.9099e-01, 1.3906e-14, 1.1327e-09],
                         [7.9096e-01, -6.1180e-01, -8.5126e-03, 1.0458e-09],
                         [-6.7348e-03, 5.2093e-03, -9.9996e-01, 4.0311e+00]])
    xyz = torch.randn(N, 3) * 2.6 - 1.3
    sc = torch.rand(N, 3) + 1
    qq = torch.rand(N, 4)
    qq = qq / qq.norm(dim=-1, keepdim=True)
    cov3d = tcc.compute_cov3d_torch_impl(sc, qq).detach()
    uv, depth = tpp.project_point_torch_impl(xyz, intr, extr, W, H, nearest=0.0, extent=1.3)
    # nearest=0 in the API; the restatement's `depth <= nearest` then also culls depth <= 0
    visible = (depth != 0).squeeze(-1)
    tep.uv = uv  # the restatement reads a module-level `uv.device`
    x1, c1 = xyz.clone().requires_grad_(), cov3d.clone().requires_grad_()
    i1, e1 = intr.clone().requires_grad_(), extr.clone().requires_grad_()
    conic, radius, tiles = tep.ewa_project_torch_impl(x1, c1, i1, e1, uv, W, H, visible)
    conic.sum().backward()
    nn = torch.nan_to_num
    np.savez_compressed(os.path.join(HERE, "ewa_project.npz"), xyz=xyz.numpy(), cov3d=cov3d.numpy(), intr=intr.numpy(),
                        extr=extr.numpy(), uv=uv.numpy(), depth=depth.numpy(), visible=visible.numpy(), W=W, H=H,
                        conic=conic.detach().numpy(), radius=radius.numpy(), tiles=tiles.numpy(),
                        dxyz=nn(x1.grad).numpy(), dcov3d=nn(c1.grad).numpy(), dintr=nn(i1.grad).numpy(),
                        dextr=nn(e1.grad).numpy())

    # compute_sh (test/test_compute_sh.py:8-13, seed 123, degree 10)
    torch.manual_seed(123)
    N = 64
    vd = torch.randn(N, 3)
    vd = (vd / vd.norm(dim=-1, keepdim=True)).double().requires_grad_()
    shc = torch.randn(N, 2, 121).double().requires_grad_()
    val = tsh.compute_sh_torch_impl(shc, vd)
    val.mean().backward()
    np.savez_compressed(os.path.join(HERE, "compute_sh.npz"), dirs=vd.detach().numpy(), shs=shc.detach().numpy(),
                        value=val.detach().numpy(), dshs=shc.grad.numpy(), ddirs=vd.grad.numpy())

    # alpha_blending (test/test_alpha_blending.py:6-63 loops, seed 121, 32x16, N=20; C reduced to 5)
    import test_alpha_blending as tab
    torch.manual_seed(121)
    w, h, bg, N, C = 32, 16, 1, 20, 5
    uv = torch.rand(N, 2)
    uv[:, 0] *= w
    uv[:, 1] *= h
    A = torch.randn(N, 2, 2)
    cv = torch.bmm(A, A.transpose(1, 2))
    conic = torch.stack([cv[:, 0, 0], cv[:, 0, 1], cv[:, 1, 1]], dim=-1)
    depth = torch.rand(N, 1) * 5
    radius = (torch.rand(N, 1) * 5).int()
    tiles = tab.get_tiles(uv, radius.squeeze(-1), w, h)
    opacity = torch.rand(N, 1)
    feature = torch.rand(N, C)
    sys.path.insert(0, ROOT)
    import oracle  # only to produce the sorted lists the loop restatement consumes
    ids, tr = oracle.sort_gaussian(uv, depth, w, h, radius, tiles)
    u1, c1 = uv.clone().requires_grad_(), conic.clone().requires_grad_()
    o1, f1 = opacity.clone().requires_grad_(), feature.clone().requires_grad_()
    img = tab.alpha_blending_torch_impl(u1, c1, o1, f1, ids, tr, bg, w, h)
    img.sum().backward()
    np.savez_compressed(os.path.join(HERE, "alpha_blending.npz"), uv=uv.numpy(), conic=conic.numpy(),
                        depth=depth.numpy(), radius=radius.numpy(), tiles=tiles.numpy(), opacity=opacity.numpy(),
                        feature=feature.numpy(), idx_sorted=ids.numpy(), tile_range=tr.numpy(), W=w, H=h, bg=bg,
                        image=img.detach().numpy(), duv=u1.grad.numpy(), dconic=c1.grad.numpy(),
                        dopacity=o1.grad.numpy(), dfeature=f1.grad.numpy())
    print("golden fixtures written to", HERE)


def gs2d_target():
    from PIL import Image
    im = Image.open(os.path.join(REF, "data", "stanford-bunny.jpg")).convert("RGB")
    im.save(os.path.join(HERE, "gs2d_target.png"), optimize=True)
    print("gs2d_target.png", im.size)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gs2d":
        gs2d_target()
        sys.exit(0)
    main()
