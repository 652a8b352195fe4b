#!/usr/bin/env python3
"""Fit a 2-D image with 3-D Gaussians -- the acceptance run for BASELINE config #2.

Same experiment as the reference's tutorial (/root/reference/tutorials/gs_2d.py:10-32 parameters and
activations, :48-64 seed / camera, :66-87 loop: Adam lr 0.01, SmoothL1 against the target, bg = 1),
written against the drop-in API so that either library can run it:

    python tutorials/gs_2d.py                       # msplat_b200, the tutorial's 512x512 bunny target, fused Adam
    python tutorials/gs_2d.py --image photo.jpg     # any RGB image (PIL)
    python tutorials/gs_2d.py --library msplat      # the reference build, if importable (torch.optim.Adam)
    python tutorials/gs_2d.py --optimizer torch     # torch.optim.Adam with msplat_b200

The default target is the tutorial's own image (data/stanford-bunny.jpg of the reference, stored losslessly as
tests/golden/gs2d_target.png); ``--image procedural`` selects a synthetic test card.  With msplat_b200 the
optimizer step is ``msplat_b200.optim.FusedAdam`` (one launch over the five parameter tensors).
No imageio / tqdm: progress goes to stdout, frames (optional) are written as PNGs with PIL.
"""
from __future__ import annotations

import argparse
import importlib
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def procedural_target(H: int, W: int) -> torch.Tensor:
    """A smooth RGB test card in [0, 1] (discs and a gradient on white): stands in for
    data/stanford-bunny.jpg, which is not redistributed with this repository."""
    y, x = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    img = torch.ones(3, H, W)
    for cx, cy, r, col in ((-0.35, -0.2, 0.38, (0.85, 0.25, 0.2)), (0.3, 0.1, 0.45, (0.2, 0.45, 0.85)),
                           (0.0, 0.45, 0.3, (0.25, 0.7, 0.3))):
        m = torch.sigmoid((r - torch.sqrt((x - cx) ** 2 + (y - cy) ** 2)) * 40.0)
        for k in range(3):
            img[k] = img[k] * (1 - m) + col[k] * m
    img *= (0.85 + 0.15 * x).clamp(0, 1)
    return img.clamp(0, 1)


BUNNY = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gs2d_target.png")


def load_target(path: str | None, size: int) -> torch.Tensor:
    if path == "procedural":
        return procedural_target(size, size)
    if path is None:
        path = BUNNY
    from PIL import Image
    import numpy as np
    im = np.asarray(Image.open(path).convert("RGB"), dtype="float32") / 255.0
    return torch.from_numpy(im).permute(2, 0, 1).contiguous()


def make_parameters(n: int, device, generator=None):
    """gs_2d.py:13-19: uniform initialisation; activations are applied in `activated`."""
    r = lambda *s: torch.rand(*s, generator=generator)
    raw = {"xyz": r(n, 3) * 2 - 1, "scale": r(n, 3), "rotate": r(n, 4), "opacity": r(n, 1), "rgb": r(n, 3)}
    return {k: torch.nn.Parameter(v.to(device)) for k, v in raw.items()}


def activated(p):
    """gs_2d.py:21-26."""
    return (p["xyz"], p["scale"].abs() + 1e-8, torch.nn.functional.normalize(p["rotate"]),
            torch.sigmoid(p["opacity"]), torch.sigmoid(p["rgb"]))


def camera(W: int, H: int, device):
    """gs_2d.py:57-64: 90 degree field of view, camera 2.5 units in front of the cloud."""
    fov = math.pi / 2.0
    fx, fy = 0.5 * W / math.tan(0.5 * fov), 0.5 * H / math.tan(0.5 * fov)
    intr = torch.tensor([fx, fy, W / 2.0, H / 2.0], dtype=torch.float32, device=device)
    extr = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 2.5]], dtype=torch.float32, device=device)
    return intr, extr


def fit(api, target: torch.Tensor, points: int, iters: int, lr: float = 0.01, seed: int = 123, log_every: int = 100,
        frames_dir: str | None = None, frame_every: int = 20, quiet: bool = False, optimizer: str = "torch",
        timing: dict | None = None):
    """Runs the optimisation; returns the list of losses (one float per iteration).  optimizer: "torch"
    (torch.optim.Adam, gs_2d.py:32) or "fused" (msplat_b200.optim.FusedAdam: the same update in one launch)."""
    device = target.device
    _, H, W = target.shape
    g = torch.Generator().manual_seed(seed)
    params = make_parameters(points, device, g)
    if optimizer == "fused":
        from msplat_b200.optim import FusedAdam
        opt = FusedAdam(list(params.values()), lr=lr)
    else:
        opt = torch.optim.Adam(list(params.values()), lr=lr)
    intr, extr = camera(W, H, device)
    loss_fn = torch.nn.SmoothL1Loss()
    losses = []
    t0 = time.time()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if timing is not None else None
    loss_dev = []
    for it in range(iters):
        if ev is not None and it == min(10, iters - 1):
            ev[0].record()  # device time of the remaining iterations (the first ones warm the allocator up)
        image = api.rasterization(*activated(params), intr, extr, W, H, 1.0)
        loss = loss_fn(image, target)
        loss.backward()
        opt.step()
        opt.zero_grad()
        if quiet and frames_dir is None:
            loss_dev.append(loss.detach())  # no host sync per iteration
            continue
        losses.append(float(loss.detach()))
        if frames_dir is not None and it % frame_every == 0:
            from PIL import Image
            os.makedirs(frames_dir, exist_ok=True)
            arr = (image.detach().clamp(0, 1).permute(1, 2, 0).cpu().numpy() * 255).astype("uint8")
            Image.fromarray(arr).save(os.path.join(frames_dir, f"frame_{it:05d}.png"))
        if not quiet and (it % log_every == 0 or it == iters - 1):
            mse = float(((image.detach().clamp(0, 1) - target) ** 2).mean())
            psnr = -10.0 * math.log10(max(mse, 1e-12))
            print(f"iter {it:5d}  loss {losses[-1]:.7f}  psnr {psnr:5.2f} dB  {(it + 1) / (time.time() - t0):7.1f} it/s",
                  flush=True)
    if ev is not None:
        ev[1].record()
        torch.cuda.synchronize()
        timing["ms_per_iteration"] = ev[0].elapsed_time(ev[1]) / max(iters - min(10, iters - 1), 1)
    if loss_dev:
        losses = [float(x) for x in torch.stack(loss_dev).cpu()]
    return losses


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--library", default="msplat_b200", help="module that provides rasterization()")
    ap.add_argument("--image", default=None, help="RGB target image; default: the tutorial's bunny; 'procedural': a test card")
    ap.add_argument("--optimizer", default=None, choices=["torch", "fused"], help="default: fused with msplat_b200")
    ap.add_argument("--size", type=int, default=512, help="side of the procedural target")
    ap.add_argument("--points", type=int, default=10000)
    ap.add_argument("--iters", type=int, default=7000)
    ap.add_argument("--frames", default=None, help="directory for PNG frames (every 20 iterations)")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("gs_2d.py needs a CUDA device: the rasterizer has no CPU path")
    api = importlib.import_module(args.library)
    target = load_target(args.image, args.size).cuda()
    opt = args.optimizer or ("fused" if args.library == "msplat_b200" else "torch")
    losses = fit(api, target, args.points, args.iters, frames_dir=args.frames, optimizer=opt)
    print(f"final loss {losses[-1]:.7f} (first {losses[0]:.7f})")


if __name__ == "__main__":
    main()
