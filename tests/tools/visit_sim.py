#!/usr/bin/env python3
"""CPU simulation of the blend kernels' list traversal (design tool, not product code).

Counts, on a scaled-down S-frustum scene with the same Gaussian density per pixel as BASELINE
config #3, how many warp-visits the backward blend makes when the unit that owns a private hit
list is (a) a whole warp covering 8x4 pixels (blend.cu today), (b) a half-warp covering 4x4 or
8x2 pixels, (c) a quarter-warp covering 4x2 pixels.  Sub-warp units of one warp advance in
lock-step over 32-entry windows of the staged batch, so a window costs max(popcount) visits.
Uses the CPU oracle for geometry, sort and ncontrib.

    python tests/tools/visit_sim.py [scale_div=3] [window=32]   (test infrastructure: it imports oracle/)
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from msplat_b200.scenes import frustum_scene  # noqa: E402


def cull_extent(cx, cy, cz, op):
    det = cx * cz - cy * cy
    tau = 2.0 * np.log(np.maximum(op * 255.0, 1e-30)) * 1.001 + 0.05
    k = tau / det
    hx = np.sqrt(np.maximum(k * cz, 0)) * 1.0001 + 0.01
    hy = np.sqrt(np.maximum(k * cx, 0)) * 1.0001 + 0.01
    dead = op * 255.0 * 1.001 < 1.0
    hx[dead] = -np.inf
    hy[dead] = -np.inf
    hs = np.sqrt(np.maximum(k * (cx + cz - 2 * cy), 0)) * 1.0001 + 0.02  # half extent of x + y
    ht = np.sqrt(np.maximum(k * (cx + cz + 2 * cy), 0)) * 1.0001 + 0.02  # half extent of x - y
    hs[dead] = -np.inf
    ht[dead] = -np.inf
    return hx, hy, hs, ht


def main():
    div = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    W, H = 1920 // div, 1080 // div
    W, H = W // 16 * 16, H // 16 * 16
    P = int(3_000_000 * (W * H) / (1920 * 1080))
    sc = frustum_scene(P, W, H, 2.0, seed=0, sh_degree=0)
    uv, depth = oracle.project_point(sc.xyz, sc.intr, sc.extr, W, H)
    vis = (depth != 0).reshape(-1)
    cov = oracle.compute_cov3d(sc.scale, sc.quat, vis)
    conic, radius, tiles = oracle.ewa_project(sc.xyz, cov, sc.intr, sc.extr, uv, W, H, vis)
    ids, tr = oracle.sort_gaussian(uv, depth, W, H, radius, tiles)
    feat = torch.rand(P, 3)
    _, final_T, ncontrib, _ = oracle.alpha_blending_forward(uv, conic, sc.opacity, feat, ids, tr, 0.0, W, H)
    uvn, cn, opn = uv.numpy(), conic.numpy(), sc.opacity.numpy().reshape(-1)
    hx, hy, hs, ht = cull_extent(cn[:, 0], cn[:, 1], cn[:, 2], opn)
    ids, tr, nc = ids.numpy(), tr.numpy(), ncontrib.numpy()
    gx, gy = W // 16, H // 16
    print(f"P={P} {W}x{H} M={ids.size} pairs={int(nc.sum())}")

    # unit shapes: name -> (units per warp, list of (x0, y0, w, h) relative to the warp's 8x4 block)
    shapes = {
        "warp 8x4": [(0, 0, 8, 4)],
        "warp 8x4 octagon": [(0, 0, 8, 4)],
        "warp 8x4 exact ellipse": [(0, 0, 8, 4)],
        "half 4x4": [(0, 0, 4, 4), (4, 0, 4, 4)],
        "half 8x2": [(0, 0, 8, 2), (0, 2, 8, 2)],
        "quarter 4x2": [(0, 0, 4, 2), (4, 0, 4, 2), (0, 2, 4, 2), (4, 2, 4, 2)],
    }
    WIN = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    visits = {k: 0 for k in shapes}
    ideal = {k: 0 for k in shapes}
    visits_nolimit = {k: 0 for k in shapes}
    useful = 0
    for ty in range(gy):
        for tx in range(gx):
            a, b = tr[ty * gx + tx]
            n = b - a
            if n <= 0:
                continue
            g = ids[a:b]
            u, v, ex, ey, es, et = uvn[g, 0], uvn[g, 1], hx[g], hy[g], hs[g], ht[g]
            qa, qb, qc = cn[g, 0], cn[g, 1], cn[g, 2]
            tau = 2.0 * np.log(np.maximum(opn[g] * 255.0, 1e-30)) * 1.001 + 0.05
            pos = np.arange(n)
            tile_nc = nc[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16]
            maxc = int(tile_nc.max())
            for w in range(8):
                bx0 = tx * 16 + (w & 1) * 8
                by0 = ty * 16 + (w >> 1) * 4
                # the backward walks positions maxc-1 .. 0 in windows of 32 aligned to maxc-1
                rpos = maxc - 1 - pos  # reverse index; window = rpos // 32 for rpos >= 0
                for name, units in shapes.items():
                    hits = []
                    for (x0, y0, uw, uh) in units:
                        X0, X1 = bx0 + x0, bx0 + x0 + uw - 1
                        Y0, Y1 = by0 + y0, by0 + y0 + uh - 1
                        umax = int(nc[Y0:Y1 + 1, X0:X1 + 1].max())
                        miss = (u + ex < X0) | (u - ex > X1) | (v + ey < Y0) | (v - ey > Y1)
                        if "exact" in name:
                            # min of q over the block: nearest point candidates on the two facing edges
                            ax0, ax1, ay0, ay1 = u - X1, u - X0, v - Y1, v - Y0
                            xe, ye = np.clip(0.0, ax0, ax1), np.clip(0.0, ay0, ay1)
                            ya = np.clip(-qb * xe / qc, ay0, ay1)
                            xb = np.clip(-qb * ye / qa, ax0, ax1)
                            q1 = qa * xe * xe + 2 * qb * xe * ya + qc * ya * ya
                            q2 = qa * xb * xb + 2 * qb * xb * ye + qc * ye * ye
                            miss = miss | (np.minimum(q1, q2) > tau)
                        if "octagon" in name:
                            sc_, tc_ = u + v, u - v
                            miss = miss | (sc_ + es < X0 + Y0) | (sc_ - es > X1 + Y1) | (tc_ + et < X0 - Y1) | (tc_ - et > X1 - Y0)
                        hits.append((~miss) & (pos < umax))
                    hm = np.stack(hits, 0)  # [units, n]
                    sel = rpos >= 0
                    win = rpos[sel] // WIN
                    nwin = int(win.max()) + 1 if win.size else 0
                    if nwin == 0:
                        continue
                    cnt = np.zeros((len(units), nwin), dtype=np.int64)
                    for k in range(len(units)):
                        np.add.at(cnt[k], win, hm[k][sel].astype(np.int64))
                    visits[name] += int(cnt.max(axis=0).sum())
                    ideal[name] += int(cnt.sum()) / len(units)
    base = visits["warp 8x4"]
    for name, vcount in visits.items():
        print(f"{name:12s} visits {vcount:10d}  x{vcount / base:.3f}   (perfectly balanced units: x{ideal[name] / base:.3f})")


if __name__ == "__main__":
    main()
