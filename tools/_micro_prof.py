import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import torch
import msplat_b200 as ms
import msplat as ref
sys.path.insert(0, os.path.join(ROOT, "tools"))
from channel_sweep import get_tiles
from torch.profiler import profile, ProfilerActivity
w = h = 800; N = 10000
for c in (4, 5):
    torch.manual_seed(121)
    uv = torch.rand([N, 2], device="cuda"); uv[:, 0] *= w; uv[:, 1] *= h
    A = torch.randn(N, 2, 2, device="cuda"); cov = torch.bmm(A, A.transpose(1, 2))
    conic = torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 1]], dim=-1)
    depth = torch.rand_like(uv[:, 0:1]) * 5
    radius = (torch.rand_like(depth) * 5).int()
    tiles = get_tiles(uv, radius.squeeze(-1), w, h).int()
    opacity = torch.rand_like(depth); feature = torch.rand([N, c], device="cuda")
    ids, tr = ms.sort_gaussian(uv, depth, w, h, radius, tiles)
    for name, api in (("ours", ms), ("ref", ref)):
        L = [t.clone().requires_grad_() for t in (uv, conic, opacity, feature)]
        def once():
            for t in L: t.grad = None
            img = api.alpha_blending(L[0], L[1], L[2], L[3], ids, tr, 0.0, w, h)
            img.sum().backward()
        for _ in range(20): once()
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(200): once()
        t_cpu = (time.time() - t0) / 200
        torch.cuda.synchronize()
        t_all = (time.time() - t0) / 200
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(20): once()
            torch.cuda.synchronize()
        ev = prof.key_averages()
        gpu = sum(e.device_time_total for e in ev) / 20
        top = sorted(ev, key=lambda e: -e.device_time_total)[:6]
        print(f"C={c} {name}: host enqueue {t_cpu*1e3:.3f} ms/iter, wall {t_all*1e3:.3f} ms/iter, GPU busy {gpu/1e3:.3f} ms/iter; " +
              "; ".join(f"{e.key[:40]} {e.device_time_total/20:.1f}us x{e.count//20}" for e in top), flush=True)
