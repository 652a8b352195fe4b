"""2-GPU probe of the data-parallel backward schedule (run under torchrun): step time vs grad_chunks."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import msplat_b200 as ms
from msplat_b200.scenes import frustum_scene, orbit_cameras
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}")); torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
P, W, H, V = 3_000_000, 1920, 1080, int(os.environ.get("VIEWS", "4"))
sc = frustum_scene(P, W, H, 2.0, seed=0, sh_degree=3).to(dev)
params = [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, sc.shs)]
extrs = torch.stack(orbit_cameras(V * world)[rank * V:(rank + 1) * V]).to(dev)
G = torch.randn(4, H, W, device=dev)
def step(chunks, sync):
    for p in params: p.grad = None
    img = ms.rasterization_sh_views(*params, sc.intr, extrs, W, H, 0.0, with_depth=True, grad_sync=sync, grad_chunks=chunks)
    (img * G).sum().backward()
    if sync is None:
        ws = [dist.all_reduce(p.grad, async_op=True) for p in params]
        for w in ws: w.wait()
res = {}
for name, chunks, sync in [("after", 1, None), ("slab1", 1, True), ("slab3", 3, True), ("slab6", 6, True), ("slab12", 12, True), ("nosync", 1, False)]:
    for _ in range(3): step(chunks, sync if sync is not False else None) if sync is not False else None
    if sync is False:
        def f():
            for p in params: p.grad = None
            img = ms.rasterization_sh_views(*params, sc.intr, extrs, W, H, 0.0, with_depth=True)
            (img * G).sum().backward()
        for _ in range(3): f()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f() if sync is False else step(chunks, sync)
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) / 5, 3)
if rank == 0: print(json.dumps(res))
dist.destroy_process_group()
