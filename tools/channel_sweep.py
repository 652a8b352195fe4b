#!/usr/bin/env python3
"""The reference's only performance artefact, reproduced for both libraries: alpha_blending fwd+bwd over the channel
count C = 1..41 at 10k Gaussians, 800x800 (/root/reference/test/test_alpha_blending_cuda.py:54-119: seed 121, uv ~ U,
conic from random 2x2 factors, radius = int(U(0,5)), loss = image.sum(), 100 iterations per channel count).

Device time per fwd+bwd (CUDA events) for the unmodified reference build and for msplat_b200 on identical tensors,
plus the image agreement, one JSON line per C and a markdown table (profiles/r2_channel_sweep.md).

    python tools/channel_sweep.py [--out gpurun_out/channel_sweep.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

import torch  # noqa: E402


def get_tiles(uv, radius, W, H):  # test_alpha_blending_cuda.py:8-44
    tl = torch.zeros_like(uv, dtype=torch.int)
    br = torch.zeros_like(uv, dtype=torch.int)
    tl[:, 0] = ((uv[:, 0] - radius) / 16)
    tl[:, 1] = ((uv[:, 1] - radius) / 16)
    br[:, 0] = ((uv[:, 0] + radius + 15) / 16)
    br[:, 1] = ((uv[:, 1] + radius + 15) / 16)
    bx, by = (W + 15) // 16, (H + 15) // 16
    tmin = torch.stack([tl[:, 0].clamp(0, bx), tl[:, 1].clamp(0, by)], -1)
    tmax = torch.stack([br[:, 0].clamp(0, bx), br[:, 1].clamp(0, by)], -1)
    d = tmax - tmin
    return d[:, 0] * d[:, 1]


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "channel_sweep.json"))
    ap.add_argument("--iters", type=int, default=100)
    args = ap.parse_args()
    import msplat_b200 as ms
    try:
        import msplat as ref
    except Exception:
        ref = None
    w = h = 800
    N = 10000
    rows = []
    for c in range(1, 42):
        torch.manual_seed(121)
        uv = torch.rand([N, 2], device="cuda")
        uv[:, 0] *= w
        uv[:, 1] *= h
        A = torch.randn(N, 2, 2, device="cuda")
        cov = torch.bmm(A, A.transpose(1, 2))
        conic = torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 1]], dim=-1)
        depth = torch.rand_like(uv[:, 0:1]) * 5
        radius = (torch.rand_like(depth) * 5).int()
        tiles = get_tiles(uv, radius.squeeze(-1), w, h).int()
        opacity = torch.rand_like(depth)
        feature = torch.rand([N, c], device="cuda")
        ids, tr = ms.sort_gaussian(uv, depth, w, h, radius, tiles)
        row = {"C": c, "keys": int(ids.numel())}
        imgs = {}
        for name, api in (("ours", ms), ("reference", ref)):
            if api is None:
                continue
            L = [t.clone().requires_grad_() for t in (uv, conic, opacity, feature)]

            def once():
                for t in L:
                    t.grad = None
                img = api.alpha_blending(L[0], L[1], L[2], L[3], ids, tr, 0.0, w, h)
                img.sum().backward()
                return img

            imgs[name] = once().detach()
            row[name + "_ms"] = round(timed(once, args.iters), 4)
        if ref is not None:
            row["speedup"] = round(row["reference_ms"] / row["ours_ms"], 2)
            row["image_max_abs_diff"] = float((imgs["ours"] - imgs["reference"]).abs().max())
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)
    with open(args.out.replace(".json", ".md"), "w") as f:
        f.write("| C | ours ms | reference ms | speed-up | image max abs diff |\n|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r['C']} | {r['ours_ms']} | {r.get('reference_ms', '-')} | {r.get('speedup', '-')} | "
                    f"{r.get('image_max_abs_diff', '-')} |\n")


if __name__ == "__main__":
    main()
