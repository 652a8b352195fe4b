"""Fused SH render path: ``rasterization_sh`` (one camera) and ``rasterization_sh_views`` (a batch
of cameras over the same Gaussians).

The reference only offers this as a chain of steps plus torch glue
(/root/reference/msplat/__init__.py:70-93 preceded by ``compute_sh``; tutorials/gs_3d.py-style):

    uv, depth = project_point(xyz, intr, extr, W, H);  visible = depth != 0
    dirs = normalize(xyz - camera_centre);  rgb = clamp_min(compute_sh(shs, dirs, visible) + 0.5, 0)
    feature = cat(rgb, depth) [optional];  cov3d = compute_cov3d(...);  conic, radius, tiles = ewa_project(...)
    ids, tile_range = sort_gaussian(...);  image = alpha_blending(...)

Here the whole per-Gaussian part is ONE forward and ONE backward kernel (csrc/render.cu) that read
the parameters once, never materialise cov3d / dirs / rgb and write straight into the blend
kernels' packed layout; sort and blend are the same kernels as the steps API.  uv, depth, radius,
tiles, idx_sorted and tile_range are bit-identical to the steps pipeline; images agree to FP32
rounding of the view-direction normalisation (tests/test_gpu_parity.py::test_render_sh_*).

A view batch shares one host sync (all M read-backs at once) and accumulates the per-Gaussian
gradients of its views inside the backward kernel (no per-view add passes): SURVEY 8f ranks 1+3.
"""
from __future__ import annotations

from typing import Tuple

import os

import torch
from torch import Tensor

from . import _lib
from ._lib import as_f32, ptr

__all__ = ["rasterization_sh", "rasterization_sh_views"]

# Two-stream schedule for view batches: the sort of view b+1 (latency-bound integer passes) runs on a
# side stream under the forward blend of view b (issue-bound), and the fused preprocess backward of
# view b (HBM-bound) under the backward blend of view b+1.  Results are unaffected (same kernels,
# same accumulation order); set to False to serialise everything on the caller's stream (used by
# bench.py for per-kernel timings).
OVERLAP = True
_side_streams = {}


SORT_STREAMS = int(os.environ.get("MSB_SORT_STREAMS", "1"))  # side streams the sorts of a view batch rotate over


def _side_stream(dev, k: int = 0) -> "torch.cuda.Stream":
    idx = torch.device(dev).index
    if idx is None:
        idx = torch.cuda.current_device()
    if (idx, k) not in _side_streams:
        # high priority: its CTAs are dispatched as soon as blend CTAs retire instead of queueing behind the
        # thousands of tile CTAs of the blend kernel launched before them
        _side_streams[(idx, k)] = torch.cuda.Stream(device=idx, priority=-1)
    return _side_streams[(idx, k)]


def rasterization_sh(
    xyz: Tensor, scale: Tensor, rotate: Tensor, opacity: Tensor, shs: Tensor, intr: Tensor, extr: Tensor,
    W: int, H: int, bg: float, *, sh_bias: float = 0.5, clamp: bool = True, with_depth: bool = False,
    nearest: float = 0.0, extent: float = 1.3, ndc: Tensor = None, return_aux: bool = False,
):
    """One camera.  xyz [P,3], scale [P,3], rotate [P,4] (r,x,y,z; not normalised inside),
    opacity [P,1], shs [P,Cs,D] with D=(deg+1)^2, intr [4], extr [3,4]|[4,4] -> image [C,H,W],
    C = Cs (+1 depth channel if ``with_depth``).  ``ndc`` [P,2] / ``return_aux``: see
    :func:`rasterization_sh_views` (aux tensors come back without the view dimension)."""
    out = rasterization_sh_views(xyz, scale, rotate, opacity, shs, intr[None], extr[None], W, H, bg, sh_bias=sh_bias,
                                 clamp=clamp, with_depth=with_depth, nearest=nearest, extent=extent,
                                 ndc=None if ndc is None else ndc[None], return_aux=return_aux)
    if return_aux:
        return out[0][0], out[1][0], out[2][0]
    return out[0]


def rasterization_sh_views(
    xyz: Tensor, scale: Tensor, rotate: Tensor, opacity: Tensor, shs: Tensor, intrs: Tensor, extrs: Tensor,
    W: int, H: int, bg: float, *, sh_bias: float = 0.5, clamp: bool = True, with_depth: bool = False,
    nearest: float = 0.0, extent: float = 1.3, grad_sync=None, grad_chunks: int = 3, ndc: Tensor = None,
    return_aux: bool = False,
):
    """B cameras over the same Gaussians.  intrs [B,4] (or [4], shared), extrs [B,3,4]|[B,4,4]
    -> images [B,C,H,W].

    Side outputs a 3DGS trainer reads every step, at no extra pass:
    ``ndc`` [B,P,2] is a dummy input whose ``.grad`` receives the screen-space gradient
    ``dL_duv * [0.5 W, 0.5 H]`` of every view (the reference's hook, msplat/alpha_blending.py:107-110,
    which densification heuristics accumulate); ``return_aux=True`` returns
    ``(images, radii [B,P] int32, visible [B,P] bool)`` with ``visible = radii > 0`` (the Gaussians
    that take part in a view, src/sort_gaussian.cu:26).

    ``grad_sync`` (view-batch data parallelism): a ``torch.distributed`` process group
    (or ``True`` for the default group).  The backward pass then returns the per-Gaussian gradients
    already SUMMED over the ranks of that group: the Gaussians are processed in ``grad_chunks``
    slabs, and the all-reduce of a finished slab (NCCL, its own stream) runs under the
    preprocess-backward kernels of the following slabs instead of after the whole backward.
    Camera gradients stay local (every rank has its own cameras).  A callable is accepted as a
    custom reducer: it is called with every finished gradient slab (in place) and may return an
    object with ``.wait()``."""
    if intrs.dim() == 1:
        intrs = intrs[None].expand(extrs.shape[0], 4)
    if ndc is not None and tuple(ndc.shape) != (extrs.shape[0], xyz.shape[0], 2):
        raise RuntimeError("rasterization_sh_views: ndc must be [B, P, 2]")
    images, radii = _RenderSHViews.apply(xyz, scale, rotate, opacity, shs, intrs, extrs, int(W), int(H), float(bg),
                                         float(sh_bias), bool(clamp), bool(with_depth), float(nearest), float(extent),
                                         grad_sync, int(grad_chunks), ndc, bool(return_aux))
    if return_aux:
        return images, radii, radii > 0
    return images


def _resolve_group(grad_sync):
    """-> process group to all-reduce over, or None when there is nothing to do."""
    import torch.distributed as dist
    if callable(grad_sync):
        return grad_sync  # custom reducer: called as grad_sync(tensor_slab) -> None | object with .wait()
    if grad_sync is None or grad_sync is False or not (dist.is_available() and dist.is_initialized()):
        return None
    group = dist.group.WORLD if grad_sync is True else grad_sync
    return group if dist.get_world_size(group) > 1 else None


class _RenderSHViews(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, scale, rotate, opacity, shs, intrs, extrs, W, H, bg, sh_bias, clamp, with_depth, nearest,
                extent, grad_sync, grad_chunks, ndc, return_aux):
        x, s, q = as_f32(xyz, "xyz"), as_f32(scale, "scale"), as_f32(rotate, "rotate")
        o, sh = as_f32(opacity, "opacity"), as_f32(shs, "shs")
        I, E = as_f32(intrs, "intrs"), as_f32(extrs, "extrs")
        P = x.shape[0]
        if x.shape != (P, 3) or s.shape != (P, 3) or q.shape != (P, 4) or o.numel() != P or sh.dim() != 3 \
                or sh.shape[0] != P:
            raise RuntimeError("rasterization_sh: xyz [P,3], scale [P,3], rotate [P,4], opacity [P,1], shs [P,Cs,D]")
        Cs, D = int(sh.shape[1]), int(sh.shape[2])
        deg = int(round(D ** 0.5)) - 1
        if (deg + 1) ** 2 != D or not 0 <= deg <= 10:
            raise RuntimeError(f"shs last dim must be (deg+1)^2 with deg <= 10, got {D}")
        if E.dim() != 3 or E.shape[1] not in (3, 4) or E.shape[2] != 4 or I.shape != (E.shape[0], 4):
            raise RuntimeError("rasterization_sh_views: intrs [B,4], extrs [B,3,4] or [B,4,4]")
        B = E.shape[0]
        C = Cs + (1 if with_depth else 0)
        dev = x.device
        L = _lib.lib()
        cpad = L.msb_blend_cpad(C)
        T = ((W + 15) // 16) * ((H + 15) // 16)
        f32, i32 = torch.float32, torch.int32
        images = torch.empty((B, C, H, W), dtype=f32, device=dev)
        views = []
        with torch.cuda.device(dev):
            totals = _lib.pinned_i64(dev, B)
            total_dev = torch.empty((B,), dtype=torch.int64, device=dev)
            # phase A: per-Gaussian preprocess + tile-count scan of every view, then ONE host sync
            for b in range(B):
                rec = torch.empty((P, 8), dtype=f32, device=dev)
                featp = torch.empty((P, cpad), dtype=f32, device=dev)
                uv = torch.empty((P, 2), dtype=f32, device=dev)
                depth = torch.empty((P,), dtype=f32, device=dev)
                radius = torch.empty((P,), dtype=i32, device=dev)
                tiles = torch.empty((P,), dtype=i32, device=dev)
                # M = sum(tiles) of view b is accumulated by the same kernel into total_dev[b]
                _lib.call("render_preprocess_forward", 1 if P else 0, L.msb_render_preprocess_fwd, dev, ptr(x), ptr(s),
                          ptr(q), ptr(o), ptr(sh), ptr(I[b]), ptr(E[b]), P, Cs, D, int(with_depth), W, H, nearest,
                          extent, sh_bias, int(clamp), ptr(rec), ptr(featp), ptr(uv), ptr(depth), ptr(radius),
                          ptr(tiles), ptr(total_dev[b:]), None)
                views.append([rec, featp, uv, depth, radius, tiles])
            totals[:B].copy_(total_dev, non_blocking=True)  # one device->host copy for the whole batch
            main = torch.cuda.current_stream(dev)
            main.synchronize()
            Ms = [int(totals[b]) for b in range(B)]
            sides = [_side_stream(dev, k) for k in range(max(1, SORT_STREAMS))] if (OVERLAP and B > 1) else None
            if sides is not None:
                for st_ in sides[1:]:
                    st_.wait_stream(main)  # the per-view tensors were produced on `main`
            # phase B: sort (side stream when overlapping) + blend (caller's stream) per view
            saved, keep = [], []
            radii = torch.stack([v[4] for v in views]) if return_aux else torch.empty((0,), dtype=i32, device=dev)
            for b in range(B):
                rec, featp, uv, depth, radius, tiles = views[b]
                M = Ms[b]
                if M >= 2 ** 30:
                    raise RuntimeError(f"rasterization_sh: {M} tile intersections exceed the supported 2^30")
                ids = torch.empty((M,), dtype=i32, device=dev)
                tr = torch.empty((T, 2), dtype=i32, device=dev)
                ws2 = torch.empty((L.msb_sort_workspace_bytes(P, M, W, H),), dtype=torch.uint8, device=dev)
                final_T = torch.empty((H, W), dtype=f32, device=dev)
                ncontrib = torch.empty((H, W), dtype=i32, device=dev)
                nk = 4 + L.msb_sort_num_passes(W, H) if (M > 0 and P > 0) else 0  # keygen, offsets, duplicate, ranges + passes
                side = sides[b % len(sides)] if sides is not None else None
                with torch.cuda.stream(side if side is not None else main):
                    _lib.call("sort_gaussian", nk, L.msb_sort_gaussian, dev, ptr(uv), ptr(depth), ptr(radius),
                              ptr(tiles), P, M, W, H, ptr(ids), ptr(tr), ptr(ws2), ws2.numel(), _lib.sm_count(dev))
                    if side is not None:
                        main.wait_event(side.record_event())
                _lib.call("blend_forward", _blend_passes_fwd(cpad, C), L.msb_blend_packed_fwd, dev, ptr(rec), ptr(featp),
                          ptr(ids), ptr(tr), bg, C, W, H, ptr(images[b]), ptr(final_T), ptr(ncontrib))
                saved += [rec, featp, tiles, ids, tr, final_T, ncontrib]
                keep += [uv, depth, radius, ws2]  # alive until both streams are joined (allocated on `main`)
                views[b] = None
            # all side-stream work is ordered before the last blend, hence before anything the caller enqueues next
            del keep
        ctx.cfg = (B, P, Cs, D, C, cpad, W, H, bg, sh_bias, clamp, with_depth)
        ctx.cam_grad = (intrs.requires_grad, extrs.requires_grad)
        ctx.grad_sync = (grad_sync, grad_chunks)
        ctx.shapes = (tuple(opacity.shape), tuple(intrs.shape), tuple(extrs.shape))
        ctx.has_ndc = ndc is not None
        ctx.save_for_backward(x, s, q, sh, I, E, *saved)
        ctx.mark_non_differentiable(radii)
        return images, radii

    @staticmethod
    def backward(ctx, dL_dimages, _dL_dradii=None):
        B, P, Cs, D, C, cpad, W, H, bg, sh_bias, clamp, with_depth = ctx.cfg
        x, s, q, sh, I, E = ctx.saved_tensors[:6]
        saved = ctx.saved_tensors[6:]
        g = as_f32(dL_dimages, "dL_dimages")
        dev = x.device
        L = _lib.lib()
        f32 = torch.float32
        dxyz = torch.empty((P, 3), dtype=f32, device=dev)
        dscale = torch.empty((P, 3), dtype=f32, device=dev)
        dquat = torch.empty((P, 4), dtype=f32, device=dev)
        dop = torch.empty((P,), dtype=f32, device=dev)
        dshs = torch.empty_like(sh)
        need_i, need_e = ctx.cam_grad
        # screen-space gradient hook: dL_duv of view b is columns 0:2 of its packed gradient record
        dndc = torch.zeros((B, P, 2), dtype=f32, device=dev) if ctx.has_ndc else None
        ndc_scale = torch.tensor([0.5 * W, 0.5 * H], dtype=f32, device=dev) if ctx.has_ndc else None
        dintr = torch.zeros((B, 4), dtype=f32, device=dev) if need_i else None
        dextr = torch.zeros((B,) + tuple(E.shape[1:]), dtype=f32, device=dev) if need_e else None
        if P == 0 or C == 0:
            for t in (dxyz, dscale, dquat, dop, dshs):
                t.zero_()
        else:
            group = _resolve_group(ctx.grad_sync[0])
            outs = (dxyz, dscale, dquat, dop, dshs)

            def pre_bwd(b, gr, gf, lo, hi):
                """fused preprocess backward of view b for the Gaussians [lo, hi) (accumulates for b > 0)"""
                _lib.call("render_preprocess_backward", 1, L.msb_render_preprocess_bwd, dev, ptr(x[lo:hi]),
                          ptr(s[lo:hi]), ptr(q[lo:hi]), ptr(sh[lo:hi]), ptr(I[b]), ptr(E[b]), ptr(saved[7 * b + 2][lo:hi]),
                          ptr(gr[lo:hi]), ptr(gf[lo:hi]), hi - lo, Cs, D, int(with_depth), sh_bias, int(clamp),
                          1 if b > 0 else 0, ptr(dxyz[lo:hi]), ptr(dscale[lo:hi]), ptr(dquat[lo:hi]), ptr(dop[lo:hi]),
                          ptr(dshs[lo:hi]), ptr(dintr[b]) if need_i else None, ptr(dextr[b]) if need_e else None)

            def blend_bwd(b, gr, gf, already_zero=False):
                rec, featp, tiles, ids, tr, final_T, ncontrib = saved[7 * b:7 * b + 7]
                _lib.call("blend_backward", _blend_passes_bwd(cpad), L.msb_blend_packed_bwd, dev, ptr(rec),
                          ptr(featp), ptr(ids), ptr(tr), bg, P, C, W, H, ptr(final_T), ptr(ncontrib), ptr(g[b]),
                          ptr(gr), ptr(gf), int(already_zero))

            with torch.cuda.device(dev):
                main = torch.cuda.current_stream(dev)
                if group is not None:
                    # data-parallel schedule: all blend backwards first (packed gradients kept per view), then
                    # the preprocess backward slab by slab; a finished slab is all-reduced while the next runs
                    import torch.distributed as dist
                    grec = [torch.empty((P, 8), dtype=f32, device=dev) for _ in range(B)]
                    gfeat = [torch.empty((P, cpad), dtype=f32, device=dev) for _ in range(B)]
                    # the packed gradient buffers are cleared on the side stream, under the backward blends of
                    # the views before (only the first blend waits for its clear)
                    side = _side_stream(dev) if OVERLAP else None
                    cleared = [None] * B
                    if side is not None:
                        side.wait_stream(main)
                        with torch.cuda.stream(side):
                            for b in range(B):
                                grec[b].zero_()
                                gfeat[b].zero_()
                                cleared[b] = side.record_event()
                    for b in range(B):
                        if side is not None:
                            main.wait_event(cleared[b])
                        blend_bwd(b, grec[b], gfeat[b], already_zero=side is not None)
                        if dndc is not None:
                            torch.mul(grec[b][:, :2], ndc_scale, out=dndc[b])
                    nchunk = max(1, min(int(ctx.grad_sync[1]), (P + 255) // 256))
                    step = ((P + nchunk - 1) // nchunk + 255) // 256 * 256  # slab starts stay 16-byte aligned
                    works = []
                    for lo in range(0, P, step):
                        hi = min(P, lo + step)
                        for b in range(B):
                            pre_bwd(b, grec[b], gfeat[b], lo, hi)
                        for t in outs:
                            if callable(group):
                                works.append(group(t[lo:hi]))
                            else:
                                works.append(dist.all_reduce(t[lo:hi], op=dist.ReduceOp.SUM, group=group,
                                                             async_op=True))
                    for w in works:
                        if w is not None:
                            w.wait()
                else:
                    side = _side_stream(dev) if (OVERLAP and B > 1) else None
                    nbuf = 2 if side is not None else 1
                    grec = [torch.empty((P, 8), dtype=f32, device=dev) for _ in range(nbuf)]
                    gfeat = [torch.empty((P, cpad), dtype=f32, device=dev) for _ in range(nbuf)]
                    done = [None] * B
                    if side is not None:
                        side.wait_stream(main)  # the output tensors were allocated (and maybe recycled) on `main`
                    for b in range(B):
                        k = b % nbuf
                        if side is not None and b >= nbuf:
                            main.wait_event(done[b - nbuf])  # the packed-gradient buffer is free again
                        # from the third view on the buffer was cleared on the side stream (below), under the
                        # backward blend of the view before: the 144 MB memset leaves the critical path
                        blend_bwd(b, grec[k], gfeat[k], already_zero=(side is not None and b >= nbuf))
                        if dndc is not None:  # on `main`, before blend_bwd(b + nbuf) rewrites the buffer
                            torch.mul(grec[k][:, :2], ndc_scale, out=dndc[b])
                        with torch.cuda.stream(side if side is not None else main):
                            if side is not None:
                                side.wait_event(main.record_event())
                            pre_bwd(b, grec[k], gfeat[k], 0, P)
                            if side is not None:
                                if b + nbuf < B:
                                    grec[k].zero_()
                                    gfeat[k].zero_()
                                done[b] = side.record_event()
                    if side is not None:
                        main.wait_stream(side)
        op_shape = ctx.shapes[0]
        return (dxyz, dscale, dquat, dop.reshape(op_shape), dshs, dintr, dextr, None, None, None, None, None, None,
                None, None, None, None, dndc, None)


def _blend_passes_fwd(cpad: int, C: int) -> int:
    return 1 if C == 0 else (cpad // 32 + (1 if cpad % 32 else 0))


def _blend_passes_bwd(cpad: int) -> int:
    return (cpad + 15) // 16 if cpad > 8 else 1
