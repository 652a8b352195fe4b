#!/usr/bin/env python3
"""bench.py -- headline benchmark: fwd+bwd renders/s @ 3M Gaussians, 1920x1080, SH degree 3,
RGB+depth (BASELINE.json configs[2], SURVEY 8d "Config #3").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|oracle]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One STEP = `views` (default 8) fwd+bwd renders of the same resident 3M-Gaussian cloud from
different cameras, loss = sum(image * G), gradients to xyz, scale, rotation, opacity and SH
coefficients summed over the views.  Two ways to write that step against the public API:
  --api steps  the chain the reference offers (and the only one it has):
                 project_point -> compute_sh(+0.5, clamp) -> cat(rgb, depth) -> compute_cov3d ->
                 ewa_project -> sort_gaussian -> alpha_blending -> backward, once per view
  --api fused  msplat_b200.rasterization_sh_views: the same maths as ONE autograd Function over
               the view batch (fused per-Gaussian kernels, gradients accumulated in-kernel);
               default for --impl ours, which also reports the steps-API number as `steps_api`
With N > 1 every rank renders its own `views` cameras (weak scaling) and the per-Gaussian
gradients are summed over the ranks once per step (NCCL; only the rows some rank touched travel).
value = N * views / step_time.

--config 5        BASELINE configs[4] instead of configs[2]: 6M Gaussians, 3840x2160, a batch of 64
                  cameras STRONG-scaled over the N ranks (64 / N views per rank), one gradient exchange
                  per step; value = 64 / step_time, "scaling": "strong".

--impl ours       msplat_b200 (default)
--impl reference  the UNMODIFIED reference CUDA build from baseline/_ref through its own public
                  API on the same tensors/config (the comparator the north star names); if that
                  build is absent, the CPU oracle port on a bounded sample (rank 0 only)
--impl oracle     the CPU oracle port on a bounded sample (what `cpu_baseline` reports)
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "fwd+bwd renders/s @3M Gaussians 1080p SH3"
P_FULL, W_FULL, H_FULL, SH_DEG, SIGMA = 3_000_000, 1920, 1080, 3, 2.0
METRIC5 = "fwd+bwd renders/s, view-batch training step @6M Gaussians 4K, 64 cameras"
P5, W5, H5, SIGMA5, VIEWS5 = 6_000_000, 3840, 2160, 3.0, 64


# ------------------------------------------------------------------------------------------------
# the workload (identical code for our library and for the reference build)
# ------------------------------------------------------------------------------------------------
def render_once(api, params, cam, W, H, G):
    xyz, scale, quat, opacity, shs = params
    intr, extr, center = cam
    uv, depth = api.project_point(xyz, intr, extr, W, H)
    visible = depth != 0
    dirs = xyz - center
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    rgb = torch.clamp_min(api.compute_sh(shs, dirs, visible.squeeze(-1)) + 0.5, 0.0)
    feature = torch.cat([rgb, depth], dim=-1)
    cov3d = api.compute_cov3d(scale, quat, visible)
    conic, radius, tiles = api.ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
    ids, tile_range = api.sort_gaussian(uv, depth, W, H, radius, tiles)
    image = api.alpha_blending(uv, conic, opacity, feature, ids, tile_range, 0.0, W, H)
    loss = (image * resolve(G)).sum()
    loss.backward()
    return loss.detach()


class Staged:
    """A step input whose host->device copy was issued on a copy stream: the consumer waits for the
    copy right before the first use (the cotangent is first needed after the first forward)."""

    def __init__(self, tensor, event):
        self.tensor, self.event = tensor, event

    def get(self):
        if self.event is not None:
            torch.cuda.current_stream().wait_event(self.event)
            self.event = None
        return self.tensor


def resolve(x):
    return x.get() if isinstance(x, Staged) else x


def make_cameras(scene, n, device):
    from msplat_b200.scenes import orbit_cameras
    cams = []
    for extr in orbit_cameras(max(n, 2), yaw_deg=20.0, shift=1.0)[:n]:
        R, t = extr[:3, :3], extr[:3, 3]
        center = -(R.T @ t)
        cams.append((scene.intr.detach().cpu().clone(), extr.clone(), center.clone()))
    return cams


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def physical_gpu_index(local):
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    ids = [x for x in cvd.split(",") if x.strip()]
    if ids and local < len(ids) and ids[local].strip().isdigit():
        return int(ids[local])
    return local


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


# algorithmic HBM bytes per STEP (V views of one rank) of each C-ABI entry point at SH `Cs x D` colours and
# `cpad` blend channels (DESIGN.md "Kernels"); the stage table divides them by the measured time per step,
# whatever number of launches (view chunks, Gaussian slabs) the step was split into
def algorithmic_bytes(name, P, M, Cs, D, cpad, V, nvis_any=None, nlive_any=None):
    nvis_any = P if nvis_any is None else nvis_any      # Gaussians touching a tile in at least one view
    nlive_any = nvis_any if nlive_any is None else nlive_any  # ... that receive a colour gradient in at least one view
    sh = 4 * Cs * D
    per_view_out = 32 + 4 * cpad + 8 + 4 + 4 + 4        # rec, featp, uv, depth, radius, tiles
    T = {
        # parameters + SH rows once, per-view packed records out
        "render_preprocess_forward": P * 44 + nvis_any * sh + V * P * per_view_out,
        # parameters once, per view tiles + packed gradients in, SH rows of the live Gaussians once,
        # geometry gradients and dL_dshs written once
        "render_preprocess_backward": P * 40 + V * P * (4 + 32 + 4 * cpad) + nlive_any * sh + P * 44 + P * sh,
        "project_point_forward": V * P * (12 + 12),
        "project_point_backward": V * P * (12 + 4 + 12 + 12),
        "compute_cov3d_forward": V * P * (12 + 16 + 1 + 24),
        "compute_cov3d_backward": V * P * (12 + 16 + 1 + 24 + 28),
        "ewa_project_forward": V * P * (12 + 24 + 8 + 1 + 12 + 8),
        "ewa_project_backward": V * P * (12 + 24 + 4 + 12 + 12 + 24),
        "compute_sh_forward": V * P * (4 * Cs * D + 12 + 1 + 4 * Cs),
        "compute_sh_backward": V * P * (2 * 4 * Cs * D + 12 + 1 + 4 * Cs + 12),
        "sort_scan": V * P * (4 + 4 + 4),
        # 20 B per Gaussian and view, 12 + 8 + 24 p + 8 B per key with p = 6 digit passes
        "sort_gaussian": V * P * 20 + M * (12 + 24 * 6 + 8),
    }
    return T.get(name)


def barrier_sync(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run_gpu(args, api, impl):
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        # NCCL kernels on a high-priority stream: their few CTAs are dispatched between the CTAs of the
        # preprocess-backward slabs they overlap with (measured at N=2: 13.5 -> 13.1 ms per step)
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from msplat_b200.parallel import FlatGrads
    from msplat_b200.scenes import frustum_scene

    ours = impl == "ours"
    P, W, H = args.gaussians, args.width, args.height
    cfg5 = args.config == 5
    if cfg5:
        if VIEWS5 % world:
            raise SystemExit(f"--config 5 shards {VIEWS5} cameras: N must divide it")
        args.views = VIEWS5 // world  # strong scaling: the 64-camera batch is split over the ranks
    scene = frustum_scene(P, W, H, args.sigma, seed=0, sh_degree=SH_DEG).to(dev)
    params = [t.clone().requires_grad_() for t in (scene.xyz, scene.scale, scene.quat, scene.opacity, scene.shs)]
    V = args.views
    cams_host = make_cameras(scene, VIEWS5 if cfg5 else V * world, "cpu")[rank * V:(rank + 1) * V]
    C = 4
    G_host = torch.randn(C, H, W, generator=torch.Generator().manual_seed(1)).pin_memory()
    G = G_host.to(dev)
    # host-side (pinned) step inputs: cameras as [V,4] / [V,3,4] / [V,3] + the cotangent
    intr_h = torch.stack([c[0] for c in cams_host]).pin_memory()
    extr_h = torch.stack([c[1] for c in cams_host]).pin_memory()
    cent_h = torch.stack([c[2] for c in cams_host]).pin_memory()
    h2d = (intr_h.numel() + extr_h.numel() + cent_h.numel() + G_host.numel()) * 4
    flat = FlatGrads(params)

    def step_steps(intrs, extrs, cents, G_):
        """the reference-style chain, one backward per view, grads accumulated by autograd"""
        flat.attach()
        flat.zero_()
        total = None
        for k in range(V):
            l = render_once(api, params, (intrs[k], extrs[k], cents[k]), W, H, G_)
            total = l if total is None else total + l
        flat.all_reduce()
        return total

    def step_fused(intrs, extrs, cents, G_):
        """one autograd Function over the view batch; gradients come back already summed"""
        for p_ in params:
            p_.grad = None
        images = api.rasterization_sh_views(*params, intrs, extrs, W, H, 0.0, with_depth=True,
                                            grad_sync=(world > 1),  # grads come back summed over the ranks
                                            grad_chunks=args.grad_chunks, view_chunk=args.view_chunk)
        loss = (images * resolve(G_)).sum()
        loss.backward()
        return loss.detach()

    def measure(step, stages=False):
        """-> (ms per step resident, renders/s resident, renders/s e2e, launches, timing)"""
        dev_in = (intr_h.to(dev), extr_h.to(dev), cent_h.to(dev), G)
        # N > 1: the caching allocator needs a few more steps than a single-GPU run to reach its steady state (blocks
        # that NCCL's stream still uses cannot be recycled at once; until enough exist every step pays cudaMalloc):
        # measured at N = 2, 3 warm-up steps leave the first timed steps 30-40 % slow.  Warm-up is untimed.
        for _ in range(max(args.warmup, 8) if world > 1 else args.warmup):
            step(*dev_in)
        barrier_sync(world)
        if ours:
            _lib.reset_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(args.steps):
            step(*dev_in)
        e1.record()
        barrier_sync(world)
        t1 = time.time()
        launches = _lib.launches() if ours else 0
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        timing = None
        if ours and stages:
            # per-C-ABI-call durations: same steps again, serialised on one stream (no sort/blend overlap),
            # each call bracketed by CUDA events on its launching stream
            from msplat_b200 import render as _render
            with _render.serialised():
                step(*dev_in)
                barrier_sync(world)
                _lib.TIMING = []
                e0.record()
                for _ in range(args.steps):
                    step(*dev_in)
                e1.record()
                barrier_sync(world)
                timing, _lib.TIMING = _lib.TIMING, None
            serial_ms = e0.elapsed_time(e1) / args.steps
            # the same per-call events with the two-stream schedule on: how long each call takes while it
            # shares the SMs with the other stream's kernels
            step(*dev_in)
            barrier_sync(world)
            _lib.TIMING = []
            step(*dev_in)
            barrier_sync(world)
            ov, _lib.TIMING = _lib.TIMING, None
            agg = {}
            for name, a, b in ov:
                d = agg.setdefault(name, [0.0, 0])
                d[0] += a.elapsed_time(b)
                d[1] += 1
            timing = (timing, serial_ms, {k: round(v[0] / v[1], 4) for k, v in agg.items()})
        # e2e: per step H2D of the step's inputs (cameras + cotangent) from pinned memory, D2H of the loss
        loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev)
        g_dev = torch.empty_like(G)
        # untimed e2e steps: the first one allocates the copy stream's buffers and the pinned staging of the loss; with
        # N > 1 the allocator needs a few steps of this loop's own allocation pattern to settle (see the warm-up above)
        for _ in range(3 if world > 1 else 1):
            with torch.cuda.stream(copy_stream):
                g_dev.copy_(G_host, non_blocking=True)
                g_ev = copy_stream.record_event()
            loss_host.copy_(step(intr_h.to(dev, non_blocking=True), extr_h.to(dev, non_blocking=True),
                                 cent_h.to(dev, non_blocking=True), Staged(g_dev, g_ev)).reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        barrier_sync(world)
        e0.record()
        # The 33 MB cotangent is copied every step on a copy stream into a reused device buffer, under the
        # forward pass that does not need it yet (the previous step's readers are done: each step ends with a
        # stream synchronisation); the cameras (a few hundred bytes) go first on the compute stream.
        for _ in range(args.steps):
            with torch.cuda.stream(copy_stream):
                g_dev.copy_(G_host, non_blocking=True)
                g_ev = copy_stream.record_event()
            ins = (intr_h.to(dev, non_blocking=True), extr_h.to(dev, non_blocking=True),
                   cent_h.to(dev, non_blocking=True), Staged(g_dev, g_ev))
            tot = step(*ins)
            loss_host.copy_(tot.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the user reads the loss every step
        e1.record()
        barrier_sync(world)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if args.trace:  # CPU + GPU timeline of three e2e steps of rank 0 (diagnostic; every rank runs the steps)
            import contextlib
            from torch.profiler import ProfilerActivity, profile
            cm = profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) if rank == 0 else contextlib.nullcontext()
            with cm as prof:
                for _ in range(3):
                    with torch.cuda.stream(copy_stream):
                        g_dev.copy_(G_host, non_blocking=True)
                        g_ev = copy_stream.record_event()
                    ins = (intr_h.to(dev, non_blocking=True), extr_h.to(dev, non_blocking=True),
                           cent_h.to(dev, non_blocking=True), Staged(g_dev, g_ev))
                    tot = step(*ins)
                    loss_host.copy_(tot.reshape(1), non_blocking=True)
                    torch.cuda.current_stream().synchronize()
            if rank == 0:
                prof.export_chrome_trace(args.trace)
            barrier_sync(world)
        n = world * V * args.steps
        return ms / args.steps, n / (ms / 1e3), n / (float(t[0]) / 1e3), launches, timing, (t0, t1)

    if ours:
        from msplat_b200 import _lib
    fused = ours and args.api == "fused"
    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    ms_per_step, value, e2e_value, launches, timing, (t_wall0, t_wall1) = measure(step_fused if fused else step_steps,
                                                                                 stages=True)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    api_note = ("msplat_b200.rasterization_sh_views (fused view-batch Function: one preprocess / sort / blend launch "
                "per view batch)" if fused else
                "steps API: project_point/compute_sh/compute_cov3d/ewa_project/sort_gaussian/alpha_blending per view")
    sync_note = ", one NCCL sum all-reduce of the per-Gaussian grads per step" if world > 1 else ""
    out = {
        "metric": METRIC5 if cfg5 else METRIC, "value": value, "unit": "renders/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if cfg5 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"S-frustum(P={P}, {W}x{H}, sigma_med={args.sigma}, seed=0), SH degree {SH_DEG}, "
                               f"RGB+depth (C=4), fwd+bwd, {V} views per rank per step"
                               + (f" ({VIEWS5} cameras over {world} ranks)" if cfg5 else "")
                               + ", grads to xyz/scale/rot/opacity/shs" + sync_note,
                   "baseline_config": 5 if cfg5 else 3,
                   "gaussians": P, "width": W, "height": H, "sh_degree": SH_DEG, "channels": C,
                   "views_per_rank_per_step": V, "parallelism": f"view-dp{world}", "api": api_note,
                   "view_chunk": args.view_chunk,
                   "grad_chunks": args.grad_chunks or (0 if world == 1 else 4 if world <= 2 else 2),
                   "cache": "inputs (~0.7 GB of parameters per render) exceed the 126 MB L2; no explicit flush"},
        "impl": impl, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "renders/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
    }
    if ours:
        out["gpu_launches"] = launches
        timing, serial_ms, overlapped = timing
        out["serial_ms_per_step"] = serial_ms  # same step with the two-stream overlap switched off (stage timings)
        out["stage_ms_two_stream"] = overlapped  # per-call durations while overlapping with the other stream
        out.update(stage_report(timing, args, api, params, cams_host, G, clocks, V, world))
        if world > 1:
            out["grad_exchange"] = {"bytes_per_step": 4 * P * (11 + 3 * (SH_DEG + 1) ** 2),
                                    "note": "dense FP32 sum all-reduce of [P, 3+3+4+1+Cs*D], one coalesced NCCL call per "
                                            "slab of Gaussians, overlapped with the preprocess backward of the next slab"}
        if fused and not args.no_steps_api and not cfg5:
            s_ms, s_val, s_e2e, _, _, _ = measure(step_steps)
            out["steps_api"] = {"value": s_val, "e2e": s_e2e, "ms_per_step": s_ms, "unit": "renders/s",
                                "note": "same workload written against the reference-style steps API of msplat_b200"}
    else:
        out["gpu_launches"] = 0
        out["cpu_baseline"] = {"value": value, "unit": "renders/s", "cores": 0, "kind": "reference-cuda",
                               "sample": "full workload on the GPU through the unmodified reference build "
                                         "(baseline/_ref); no CPU implementation exists in the reference"}
    if rank == 0:
        if ours and world == 1 and not args.no_cpu_baseline and not cfg5:
            out["cpu_baseline"] = cpu_baseline(args)
            if not args.no_other_configs:
                out["other_configs"] = other_configs(api, dev)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def batch_counts(api, params, cams, W, H, C, G):
    """Work counters of one step of this rank (all its views), from the steps API of our library:
    M (keys), traversed pairs = sum(ncontrib), blended pairs (entries passing the power / alpha
    tests before termination), Gaussians touching a tile / receiving a colour gradient in at least one view."""
    from msplat_b200 import _lib
    from msplat_b200._lib import ptr
    from msplat_b200.alpha_blending import _blend_backward, _blend_forward
    L = _lib.lib()
    P = params[0].shape[0]
    Cs = params[4].shape[1]
    M = pairs = blended = 0
    vis_any = torch.zeros(P, dtype=torch.bool, device=params[0].device)
    live_any = torch.zeros_like(vis_any)
    with torch.no_grad():
        xyz, scale, quat, opacity, shs = [p.detach() for p in params]
        feat = torch.rand(P, C, device=xyz.device)
        for cam in cams:
            intr, extr = cam[0].to(xyz.device), cam[1].to(xyz.device)
            uv, depth = api.project_point(xyz, intr, extr, W, H)
            vis = depth != 0
            cov = api.compute_cov3d(scale, quat, vis)
            conic, radius, tiles = api.ewa_project(xyz, cov, intr, extr, uv, W, H, vis)
            ids, tr = api.sort_gaussian(uv, depth, W, H, radius, tiles)
            _, final_T, ncontrib, packed = _blend_forward(uv, conic, opacity.reshape(-1, 1), feat, ids, tr, 0.0, W, H)
            pairs += int(ncontrib.sum())
            M += int(ids.numel())
            vis_any |= tiles > 0
            cnt = torch.empty((H, W), dtype=torch.int32, device=xyz.device)
            _lib.call("blend_count", 1, L.msb_blend_packed_count, xyz.device, ptr(packed), ptr(ids), ptr(tr), W, H, 1,
                      ptr(cnt))
            blended += int(cnt.sum())
            dfeat = _blend_backward(feat, ids, tr, 0.0, W, H, final_T, ncontrib, G[:C].contiguous(), packed)[3]
            live_any |= (dfeat[:, :Cs] != 0).any(dim=1)
            del dfeat, packed, ids, cnt
    return {"keys": M, "pairs": pairs, "blended_pairs": blended, "nvis_any": int(vis_any.sum()),
            "nlive_any": int(live_any.sum())}


def stage_report(timing, args, api, params, cams, G, clocks, views, world):
    """Per-C-ABI-call durations (CUDA events recorded on the launching stream inside the timed region, schedule
    serialised) summed PER STEP of this rank -> roofline of the dominant stage + a per-stage table.  All rates are
    per-step work / per-step time, so they do not depend on how many launches (view chunks, Gaussian slabs) a step
    was split into."""
    P, W, H = args.gaussians, args.width, args.height
    Cs, D, C = 3, (SH_DEG + 1) ** 2, 4
    cpad = 4
    agg = {}
    for name, a, b in timing:
        d = agg.setdefault(name, [0.0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += 1
    cnt = batch_counts(api, params, cams, W, H, C, G)
    pairs, blended, M = cnt["pairs"], cnt["blended_pairs"], cnt["keys"]
    hbm, sm_max, src = measured_peaks()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    f_hz = (clocks["sm_mhz"] if clocks else sm_max) * 1e6
    total_ms = sum(v[0] for v in agg.values())
    is_blend = lambda n: n.startswith("alpha_blending") or n.startswith("blend_")
    # dram__bytes_read.sum + dram__bytes_write.sum per STEP of each stage from the committed `ncu --set full`
    # capture of this same workload (profiles/ncu_traffic.json, not measured in this run; null for other sizes)
    traffic = {}
    if (P, W, H, views, world) == (P_FULL, W_FULL, H_FULL, 8, 1):
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            traffic = {}
    stages = {}
    for name, (tot, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        ms = tot / args.steps  # per step of this rank
        st = {"ms_per_step": round(ms, 4), "ms_per_render": round(ms / views, 4), "share": round(tot / total_ms, 4),
              "calls_per_step": n / args.steps}
        ab = algorithmic_bytes(name, P, M, Cs, D, cpad, views, cnt["nvis_any"], cnt["nlive_any"])
        if ab is not None:
            st["GBps"] = round(ab / (ms * 1e-3) / 1e9, 1)
            st["hbm_frac"] = round(st["GBps"] / hbm, 3)
            st["algorithmic_MB_per_step"] = round(ab / 1e6, 1)
        if isinstance(traffic.get(name), (int, float)):
            st["dram_traffic_MB_per_step"] = round(traffic[name] / 1e6, 1)
            st["GBps_measured_bytes"] = round(traffic[name] / (ms * 1e-3) / 1e9, 1)
            st["hbm_frac_measured_bytes"] = round(st["GBps_measured_bytes"] / hbm, 3)
        if is_blend(name):
            st["Gpairs_per_s"] = round(pairs / (ms * 1e-3) / 1e9, 2)
            st["Gblended_pairs_per_s"] = round(blended / (ms * 1e-3) / 1e9, 2)
        stages[name] = st
    dom = next(iter(stages))
    roof = {"kernel": dom}
    if is_blend(dom):
        bwd = dom.endswith("backward")
        lane_ops = (32 + 5 * C) if bwd else (13 + C)
        mufu = 2 if bwd else 1
        peak = min(sms * 128 * f_hz / lane_ops, sms * 16 * f_hz / mufu) / 1e9
        ach = stages[dom]["Gpairs_per_s"]
        ach_b = stages[dom]["Gblended_pairs_per_s"]
        roof.update({"bound": "fp32_issue", "achieved": ach, "peak": round(peak, 1), "unit": "Gpairs/s",
                     "frac": round(ach / peak, 4), "traffic": traffic.get(dom),
                     # distance to the roof on work that HAS to be executed: only pairs that pass the alpha test need
                     # the full per-pair arithmetic (the traversed-but-skipped ones are culled per warp)
                     "frac_blended_pairs": round(ach_b / peak, 4), "achieved_blended": ach_b,
                     "issue_active_ncu": traffic.get("_issue_active", {}).get(dom),
                     "note": f"SURVEY 8d: pair = sum(ncontrib) = {pairs} per step of {views} views; blended pairs (pass the "
                             f"power/alpha tests) = {blended}; peak = min(SMs*128*f/{lane_ops} lane-ops, SMs*16*f/{mufu} "
                             f"MUFU) at {sms} SMs, f = {f_hz/1e6:.0f} MHz (clock observed during the run). `frac` charges "
                             f"every traversed pair the full arithmetic and can exceed what the kernel executes (most "
                             f"traversed pairs are culled per warp); `frac_blended_pairs` charges only pairs that blend "
                             f"and is the distance to the roof (<= 1).  Not an HBM/tensor kernel: DRAM traffic is <10% "
                             f"of peak (profiles/); `traffic` and `issue_active_ncu` come from the committed ncu capture "
                             f"(profiles/ncu_traffic.json), not from this run"})
    else:
        ach = stages[dom].get("GBps", 0.0)
        roof.update({"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": round(ach / hbm, 4),
                     "traffic": traffic.get(dom), "note": f"peak: {src}"})
    # the largest HBM-bound stage in the contract's own roofline schema (the dominant stage above is issue-bound)
    hb = [(n, st) for n, st in stages.items() if "GBps" in st and not is_blend(n)]
    roof_hbm = None
    if hb:
        n, st = max(hb, key=lambda kv: kv[1]["ms_per_step"])
        roof_hbm = {"kernel": n, "bound": "hbm", "achieved": st["GBps"], "peak": hbm, "unit": "GB/s",
                    "frac": round(st["GBps"] / hbm, 4), "traffic": traffic.get(n),
                    "achieved_measured_bytes": st.get("GBps_measured_bytes"),
                    "note": f"algorithmic bytes {st['algorithmic_MB_per_step']} MB per step (DESIGN.md section 5, SURVEY "
                            f"8d) / {st['ms_per_step']} ms; peak: {src}; traffic = DRAM bytes per step from the committed "
                            f"ncu capture; achieved_measured_bytes = traffic / time"}
    # secondary: the sort (the north star asks for its achieved GB/s)
    if "sort_gaussian" in stages:
        sg = stages["sort_gaussian"]
        sg["Gkeys_per_s"] = round(M / (sg["ms_per_step"] * 1e-3) / 1e9, 3)
        sg["note"] = (f"M = {M} keys per step ({views} views, one batched sort per view chunk), 4 depth-digit passes over "
                      f"the emitting Gaussians + tile-digit passes over the keys; GBps uses SURVEY's 172 B/key model, "
                      f"GBps_measured_bytes the DRAM bytes ncu measured; peak {hbm} GB/s {src}")
    return {"roofline": roof, "roofline_hbm": roof_hbm, "stages": stages, "pairs_per_step": pairs,
            "blended_pairs_per_step": blended, "keys_per_step": M, "gaussians_touching_a_tile_any_view": cnt["nvis_any"],
            "gaussians_with_colour_gradient_any_view": cnt["nlive_any"]}


# ------------------------------------------------------------------------------------------------
# CPU oracle port (cpu_baseline / --impl oracle)
# ------------------------------------------------------------------------------------------------
def cpu_baseline(args, sample=None):
    import oracle
    from msplat_b200.scenes import frustum_scene
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P, W, H = args.gaussians, args.width, args.height
    Ps = sample or P  # the full workload: one render of all P Gaussians (about 15-40 s of CPU work)
    sc = frustum_scene(P, W, H, args.sigma, seed=0, sh_degree=SH_DEG)
    params = [t[:Ps].clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, sc.shs)]
    center = sc.cam_center
    G = torch.randn(4, H, W, generator=torch.Generator().manual_seed(1))

    class Api:
        project_point = staticmethod(oracle.project_point)
        compute_sh = staticmethod(oracle.compute_sh)
        compute_cov3d = staticmethod(lambda s, q, v: oracle.compute_cov3d(s, q, v.reshape(-1)))
        ewa_project = staticmethod(lambda x, c, i, e, uv, W, H, v: oracle.ewa_project(x, c, i, e, uv, W, H, v.reshape(-1)))
        sort_gaussian = staticmethod(oracle.sort_gaussian)
        alpha_blending = staticmethod(oracle.alpha_blending)

    oracle.steps.build_blend_ref()
    t0 = time.time()
    render_once(Api, params, (sc.intr, sc.extr, center), W, H, G)
    dt = time.time() - t0
    scale = P / Ps
    what = (f"one fwd+bwd render of the full workload ({P} Gaussians at {W}x{H})" if Ps == P else
            f"one fwd+bwd render of the first {Ps} of the {P} Gaussians at {W}x{H}, value extrapolated x{scale:.1f} "
            f"linearly in P (the blend cost is not linear in P: a sample, not a measurement of the workload)")
    return {"value": 1.0 / (dt * scale), "unit": "renders/s", "cores": cores, "kind": "port",
            "sample": f"{what} by the CPU oracle (torch + C/OpenMP blend) in {dt:.1f} s on {cores} threads"}


def other_configs(api, dev):
    """Summary numbers of the BASELINE configs that are not the bench line (bounded: a few seconds each).
    Parity of these configs is in tests/test_gpu_configs.py; tools/config_bench.py times the reference beside."""
    import oracle
    from msplat_b200.scenes import bunny2d_scene, cube_scene, frustum_scene, orbit_cameras
    res = {}

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

    try:  # 1: the CPU oracle's own case (10k Gaussians, 256x256, SH3 RGB), vectorised oracle on the host cores
        sc = cube_scene(10000, 256, 256, seed=0, sh_degree=3)
        G = torch.randn(3, 256, 256, generator=torch.Generator().manual_seed(1))
        oracle.steps.build_blend_ref()

        def cpu_once():
            L = [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, sc.shs)]
            uv, depth = oracle.project_point(L[0], sc.intr, sc.extr, 256, 256)
            vis = (depth != 0).reshape(-1)
            dirs = L[0] - sc.cam_center
            dirs = dirs / dirs.norm(dim=-1, keepdim=True)
            rgb = torch.clamp_min(oracle.compute_sh(L[4], dirs, vis) + 0.5, 0.0)
            cov = oracle.compute_cov3d(L[1], L[2], vis)
            conic, radius, tiles = oracle.ewa_project(L[0], cov, sc.intr, sc.extr, uv, 256, 256, vis)
            ids, tr = oracle.sort_gaussian(uv, depth, 256, 256, radius, tiles)
            (oracle.alpha_blending(uv, conic, L[3], rgb, ids, tr, 0.0, 256, 256) * G).sum().backward()
            return int(ids.numel())

        cpu_once()
        ts = []
        for _ in range(5):
            t0 = time.time()
            M1 = cpu_once()
            ts.append(time.time() - t0)
        scd = sc.to(dev)
        Gd = G.to(dev)

        def gpu_once():
            L = [t.clone().requires_grad_() for t in (scd.xyz, scd.scale, scd.quat, scd.opacity, scd.shs)]
            (api.rasterization_sh(*L, scd.intr, scd.extr, 256, 256, 0.0) * Gd).sum().backward()

        res["config1"] = {"what": "10k Gaussians, 256x256, SH3 RGB fwd+bwd", "keys": M1,
                          "cpu_oracle_renders_per_s": round(1.0 / statistics.median(ts), 3), "cpu_cores": os.cpu_count(),
                          "ours_ms": round(timed(gpu_once, 20, 3), 4)}
    except Exception as e:  # pragma: no cover
        res["config1"] = {"error": repr(e)[:200]}
    try:  # 2: gs_2d initialisation (sort-dominated)
        sc = bunny2d_scene(100000, 512, 512, seed=123).to(dev)
        rgb = torch.sigmoid(torch.rand(100000, 3, generator=torch.Generator().manual_seed(1))).to(dev)
        target = torch.rand(3, 512, 512, generator=torch.Generator().manual_seed(2)).to(dev)
        keys = {}

        def once2():
            L = [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, rgb)]
            img = api.rasterization(*L, sc.intr, sc.extr, 512, 512, 1.0)
            torch.nn.functional.smooth_l1_loss(img, target).backward()

        res["config2"] = {"what": "gs_2d initialisation: 100k Gaussians, 512x512 RGB, rasterization() fwd+bwd (M ~ 65.8M keys)",
                          "ours_ms": round(timed(once2, 5, 2), 4)}
        del sc, rgb, target
    except Exception as e:  # pragma: no cover
        res["config2"] = {"error": repr(e)[:200]}
    try:  # 4: SH degree 10, 32 channels, 1M Gaussians, 1080p (15.5 GB of coefficients, generated on the device)
        sc = frustum_scene(1_000_000, 1920, 1080, 2.0, seed=0, with_sh=False).to(dev)
        gen = torch.Generator(device=dev).manual_seed(3)
        shs = 0.1 * torch.randn(1_000_000, 32, 121, device=dev, generator=gen)
        shs[:, :, 0] *= 5.0
        shs.requires_grad_()
        G4 = torch.randn(32, 1080, 1920, device=dev, generator=gen)
        L4 = [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity)] + [shs]

        def once4():
            for t in L4:
                t.grad = None
            (api.rasterization_sh(*L4, sc.intr, sc.extr, 1920, 1080, 0.0) * G4).sum().backward()

        res["config4"] = {"what": "1M Gaussians, 1080p, SH degree 10, 32-channel feature map, fused fwd+bwd",
                          "sh_elements": int(shs.numel()), "ours_ms": round(timed(once4, 3, 1), 3)}
        del sc, shs, G4, L4
        torch.cuda.empty_cache()
    except Exception as e:  # pragma: no cover
        res["config4"] = {"error": repr(e)[:200]}
    try:  # 5: 6M Gaussians, 4K, 8 of the 64 cameras as one view batch (the full 64-camera step: bench.py --config 5)
        sc = frustum_scene(P5, W5, H5, SIGMA5, seed=0, sh_degree=3).to(dev)
        ex = torch.stack(orbit_cameras(VIEWS5)[:8]).to(dev)
        G5 = torch.randn(4, H5, W5, device=dev)
        L5 = [t.clone().requires_grad_() for t in (sc.xyz, sc.scale, sc.quat, sc.opacity, sc.shs)]

        def once5():
            for t in L5:
                t.grad = None
            (api.rasterization_sh_views(*L5, sc.intr, ex, W5, H5, 0.0, with_depth=True) * G5).sum().backward()

        ms5 = timed(once5, 2, 1)
        res["config5"] = {"what": "6M Gaussians, 3840x2160, SH3 RGB+depth: one batch of 8 of the 64 cameras, fwd+bwd",
                          "ours_ms_per_batch": round(ms5, 3), "ours_ms_per_view": round(ms5 / 8, 3)}
        del sc, ex, G5, L5
        torch.cuda.empty_cache()
    except Exception as e:  # pragma: no cover
        res["config5"] = {"error": repr(e)[:200]}
    return res


def run_oracle(args, as_reference=False):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args)
    out = {"metric": METRIC, "value": cb["value"], "unit": "renders/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
           "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "impl": "reference" if as_reference else "oracle",
           "config": {"workload": cb["sample"]}, "cpu_baseline": cb, "gpu_launches": 0,
           "e2e": {"value": cb["value"], "unit": "renders/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def import_reference():
    p = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(p, "msplat")):
        return None
    sys.path.insert(0, p)
    try:
        import msplat
        return msplat
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "oracle"])
    ap.add_argument("--views", type=int, default=8, help="renders per rank per step (BASELINE config 5: 64 views / 8 GPUs)")
    ap.add_argument("--gaussians", type=int, default=P_FULL)
    ap.add_argument("--width", type=int, default=W_FULL)
    ap.add_argument("--height", type=int, default=H_FULL)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--api", default="fused", choices=["fused", "steps"],
                    help="--impl ours only: fused view-batch Function (default) or the reference-style steps API")
    ap.add_argument("--no-steps-api", action="store_true", help="skip the secondary steps-API measurement")
    ap.add_argument("--grad-chunks", type=int, default=0,
                    help="N > 1: Gaussian slabs of the backward whose all-reduce overlaps the next slab's kernels "
                         "; 0 = the library's default (4 slabs for 2 ranks, 2 beyond; ms per step at N = 8 / 4 with 1, 2, 3, 4 slabs: "
                         "17.65, 17.23, -, 17.68 / -, 17.09, 17.60, 17.55)")
    ap.add_argument("--view-chunk", type=int, default=0, help="views per batched launch (0 = all views of the rank)")
    ap.add_argument("--config", type=int, default=3, choices=[3, 5],
                    help="BASELINE config: 3 = headline (default), 5 = 6M Gaussians, 4K, 64 cameras strong-scaled over N")
    ap.add_argument("--sigma", type=float, default=None)
    ap.add_argument("--no-other-configs", action="store_true", help="skip the config #1/#2/#4/#5 summary numbers")
    ap.add_argument("--trace", default=None, help="export a torch.profiler chrome trace of three e2e steps to this file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.config == 5:
        if args.gaussians == P_FULL:
            args.gaussians = P5
        if (args.width, args.height) == (W_FULL, H_FULL):
            args.width, args.height = W5, H5
        args.sigma = SIGMA5 if args.sigma is None else args.sigma
        if args.view_chunk == 0:
            args.view_chunk = 8
    args.sigma = SIGMA if args.sigma is None else args.sigma
    if args.impl == "oracle":
        return run_oracle(args)
    if args.impl == "reference":
        ref = import_reference()
        if ref is None or not torch.cuda.is_available():
            return run_oracle(args, as_reference=True)
        return run_gpu(args, ref, "reference")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: msplat_b200 has no CPU path")
    import msplat_b200
    run_gpu(args, msplat_b200, "ours")


if __name__ == "__main__":
    main()
