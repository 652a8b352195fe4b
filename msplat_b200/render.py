"""Fused SH render path: ``rasterization_sh`` (one camera) and ``rasterization_sh_views`` (a batch
of cameras over the same Gaussians).

The reference only offers this as a chain of steps plus torch glue
(/root/reference/msplat/__init__.py:70-93 preceded by ``compute_sh``; tutorials/gs_3d.py-style):

    uv, depth = project_point(xyz, intr, extr, W, H);  visible = depth != 0
    dirs = normalize(xyz - camera_centre);  rgb = clamp_min(compute_sh(shs, dirs, visible) + 0.5, 0)
    feature = cat(rgb, depth) [optional];  cov3d = compute_cov3d(...);  conic, radius, tiles = ewa_project(...)
    ids, tile_range = sort_gaussian(...);  image = alpha_blending(...)

Here a whole VIEW BATCH is one virtual scene: B views x P Gaussians are B*P virtual
Gaussians (id = view * P' + index, P' = P rounded up to 4) on a virtual grid of B*T tiles
(id = view * T + tile).  Per chunk of views (all B by default) that is

    1 launch   msb_render_preprocess_fwd_views   parameters + SH rows read once, per-view packed records out
    1 sort     msb_sort_gaussian_views           (view | tile | depth) keys, one onesweep sort
    1 launch   msb_blend_packed_fwd_views        blockIdx.z = view
    1 launch   msb_blend_packed_bwd_views
    1 launch   msb_render_preprocess_bwd_views   gradients summed over the views in registers / at the L2

instead of B times each.  uv, depth, radius, tiles, the per-view order of idx_sorted and tile_range are
bit-identical to the steps pipeline; images agree to FP32 rounding of the view-direction normalisation
(tests/test_gpu_render_sh.py).  One host sync per call (all M read-backs at once).
"""
from __future__ import annotations

import os
import threading
from contextlib import contextmanager

import torch
from torch import Tensor

from . import _lib
from ._lib import as_f32, ptr

__all__ = ["rasterization_sh", "rasterization_sh_views", "serialised"]

# Two-stream schedule when a view batch is processed in several chunks: the sort of chunk k+1
# (integer passes) runs on a side stream under the forward blend of chunk k (issue-bound), and the fused
# preprocess backward of chunk k (HBM-bound) under the backward blend of chunk k+1.  The packed gradient
# buffers are always cleared on the side stream, under the forward sort/blend.  Results do not depend on it.
# ``with serialised():`` puts everything on the caller's stream for the calling thread (per-kernel timings).
OVERLAP = True
_tls = threading.local()

# views per chunk of a batch (0 = the whole batch in one chunk); MSB_VIEW_CHUNK overrides the default
VIEW_CHUNK = int(os.environ.get("MSB_VIEW_CHUNK", "0"))
M_MAX = 2 ** 31 - 1  # int32 positions in idx_sorted, the reference's bound (msplat/sort_gaussian.py:42)


@contextmanager
def serialised():
    prev = getattr(_tls, "serial", False)
    _tls.serial = True
    try:
        yield
    finally:
        _tls.serial = prev


def _overlap() -> bool:
    return OVERLAP and not getattr(_tls, "serial", False)


def _side_stream(dev) -> "torch.cuda.Stream":
    """Per-thread, per-device side stream (high priority: its CTAs are dispatched as soon as blend CTAs
    retire instead of queueing behind the thousands of tile CTAs launched before them)."""
    idx = torch.device(dev).index
    if idx is None:
        idx = torch.cuda.current_device()
    cache = getattr(_tls, "streams", None)
    if cache is None:
        cache = _tls.streams = {}
    if idx not in cache:
        cache[idx] = torch.cuda.Stream(device=idx, priority=-1)
    return cache[idx]


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
    nearest: float = 0.0, extent: float = 1.3, grad_sync=None, grad_chunks: int = None, ndc: Tensor = None,
    return_aux: bool = False, view_chunk: int = None, stats: dict = None,
):
    """B cameras over the same Gaussians.  intrs [B,4] (or [4], shared), extrs [B,3,4]|[B,4,4]
    -> images [B,C,H,W].  ``view_chunk``: views per batched launch (default: all B).

    Side outputs a 3DGS trainer reads every step, at no extra pass:
    ``ndc`` [B,P,2] is a dummy input whose ``.grad`` receives the screen-space gradient
    ``dL_duv * [0.5 W, 0.5 H]`` of every view (the reference's hook, msplat/alpha_blending.py:107-110,
    which densification heuristics accumulate); ``return_aux=True`` returns
    ``(images, radii [B,P] int32, visible [B,P] bool)`` with ``visible = radii > 0`` (the Gaussians
    that take part in a view, src/sort_gaussian.cu:26).  ``stats`` (a dict) receives the cross-view
    accumulators densification reads: ``max_radii`` [P] int32 (forward) and, after backward,
    ``ndc_grad_norm_sum`` [P] = sum over views of ||dL_dndc|| and ``ndc_grad_count`` [P] (views in which the
    Gaussian was visible).

    The gradient w.r.t. ``extrs`` includes the dependence of the view direction on the camera centre
    (-R^T t), which a steps pipeline only has if it does not detach the centre.

    ``grad_sync`` (view-batch data parallelism): a ``torch.distributed`` process group
    (or ``True`` for the default group).  The backward pass then returns the per-Gaussian gradients
    already SUMMED over the ranks of that group: the last preprocess-backward launch is cut into
    ``grad_chunks`` slabs of Gaussians (default: 4 for two ranks, 2 beyond), and the sum all-reduce of a finished slab (one coalesced NCCL call for
    its five tensors, NCCL's own stream) runs under the kernels of the following slabs instead of after the
    whole backward.  Camera gradients stay local (every rank has its own cameras).  A callable is accepted as a
    custom reducer: it is called with every finished gradient slab (in place, sum semantics) and may return an
    object with ``.wait()``."""
    if intrs.dim() == 1:
        intrs = intrs[None].expand(extrs.shape[0], 4)
    if ndc is not None and tuple(ndc.shape) != (extrs.shape[0], xyz.shape[0], 2):
        raise RuntimeError("rasterization_sh_views: ndc must be [B, P, 2]")
    vc = VIEW_CHUNK if view_chunk is None else int(view_chunk)
    images, radii = _RenderSHViews.apply(xyz, scale, rotate, opacity, shs, intrs, extrs, int(W), int(H), float(bg),
                                         float(sh_bias), bool(clamp), bool(with_depth), float(nearest), float(extent),
                                         grad_sync, 0 if grad_chunks is None else int(grad_chunks), ndc, bool(return_aux), vc, stats)
    if return_aux:
        return images, radii, radii > 0
    return images


def _m_guess() -> dict:
    """per-thread memo: tile intersections per view chunk of the previous call with the same shape"""
    g = getattr(_tls, "m_guess", None)
    if g is None:
        g = _tls.m_guess = {}
    return g


def _resolve_group(grad_sync):
    """-> process group to all-reduce over, a custom reducer, or None when there is nothing to do."""
    import torch.distributed as dist
    if callable(grad_sync):
        return grad_sync  # custom reducer: called as grad_sync(tensor) -> None | object with .wait()
    if grad_sync is None or grad_sync is False or not (dist.is_available() and dist.is_initialized()):
        return None
    group = dist.group.WORLD if grad_sync is True else grad_sync
    return group if dist.get_world_size(group) > 1 else None


def _default_grad_chunks(group) -> int:
    """Slabs of the pipelined gradient exchange when the caller does not say.  Measured on B200 / NVSwitch, BASELINE
    config #3, ms per step: 2 ranks 16.27 (2 slabs) vs 16.10 (4); 4 ranks 17.09 (2) / 17.60 (3) / 17.55 (4);
    8 ranks 17.65 (1) / 17.23 (2) / 17.68 (4) -- the ring takes longer per call as it grows, and every extra call
    competes with the backward kernels for the SMs."""
    if callable(group):
        return 2
    import torch.distributed as dist
    return 4 if dist.get_world_size(group) <= 2 else 2


def _reduce_many(group, tensors, dev):
    """sum-reduce every tensor in place over the group, asynchronously (or through the custom reducer): the
    tensors of one slab go out as ONE coalesced NCCL group call.  -> list of objects with .wait() (or None)"""
    if callable(group):
        return [group(t) for t in tensors]
    import torch.distributed as dist
    try:
        with dist._coalescing_manager(group=group, device=torch.device(dev), async_ops=True) as cm:
            for t in tensors:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return [cm]
    except Exception:  # backends without coalescing support: one call per tensor
        return [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True) for t in tensors]


def _chunks(B, vc):
    vc = B if vc <= 0 else min(vc, B)
    return [(b0, min(vc, B - b0)) for b0 in range(0, B, vc)]


def _split_for_sort(chunks, Ms):
    """Chunks whose views hold more than 2^31 - 1 tile intersections together are split (greedily)."""
    out = []
    for b0, nb in chunks:
        start, acc = b0, 0
        for b in range(b0, b0 + nb):
            if Ms[b] > M_MAX:
                raise RuntimeError(f"rasterization_sh: view {b} has {Ms[b]} tile intersections, more than the "
                                   f"supported 2^31 - 1 (int32 positions, like the reference's int32 cumsum)")
            if acc + Ms[b] > M_MAX:
                out.append((start, b - start))
                start, acc = b, 0
            acc += Ms[b]
        out.append((start, b0 + nb - start))
    return out


class _RenderSHViews(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, scale, rotate, opacity, shs, intrs, extrs, W, H, bg, sh_bias, clamp, with_depth, nearest,
                extent, grad_sync, grad_chunks, ndc, return_aux, view_chunk, stats):
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
        es = int(E.shape[1]) * 4  # floats between consecutive extrinsics
        C = Cs + (1 if with_depth else 0)
        dev = x.device
        L = _lib.lib()
        cpad = L.msb_blend_cpad(C)
        T = ((W + 15) // 16) * ((H + 15) // 16)
        Pp = (P + 3) // 4 * 4  # rows per view of the view-major buffers: keeps every view's slabs 16-byte aligned
        chunks = _chunks(B, view_chunk)
        if max(nb for _, nb in chunks) * max(Pp, T) > M_MAX:
            raise RuntimeError("rasterization_sh_views: views per chunk x Gaussians (or tiles) must stay below 2^31; "
                               "pass a smaller view_chunk")
        f32, i32 = torch.float32, torch.int32
        need_grad = any(ctx.needs_input_grad[:7])
        overlap = _overlap()
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            rec = torch.empty((B, Pp, 8), dtype=f32, device=dev)
            featp = torch.empty((B, Pp, cpad), dtype=f32, device=dev)
            uv = torch.empty((B, Pp, 2), dtype=f32, device=dev)
            depth = torch.empty((B, Pp), dtype=f32, device=dev)
            radius = torch.empty((B, Pp), dtype=i32, device=dev)
            tiles = torch.empty((B, Pp), dtype=i32, device=dev)
            if Pp != P:  # padding rows take no part in the sort
                tiles[:, P:].zero_()
                radius[:, P:].zero_()
            totals = _lib.pinned_i64(dev, B)
            total_dev = torch.empty((B,), dtype=torch.int64, device=dev)
            # phase A: per-Gaussian preprocess of every chunk (M per view accumulated in-kernel), ONE host sync
            for b0, nb in chunks:
                _lib.call("render_preprocess_forward", 1 if P else 0, L.msb_render_preprocess_fwd_views, dev, ptr(x),
                          ptr(s), ptr(q), ptr(o), ptr(sh), ptr(I[b0]), ptr(E[b0]), es, P, nb, Pp, Cs, D,
                          int(with_depth), W, H, nearest, extent, sh_bias, int(clamp), ptr(rec[b0]), ptr(featp[b0]),
                          ptr(uv[b0]), ptr(depth[b0]), ptr(radius[b0]), ptr(tiles[b0]), ptr(total_dev[b0:]))
            totals[:B].copy_(total_dev, non_blocking=True)  # one device->host copy for the whole batch
            done = main.record_event()
            # while the device works on phase A, the host prepares everything phase B needs: output tensors, the
            # cleared packed-gradient buffers of the backward blend, and (sized from the previous call with the
            # same shape) the sort's outputs and workspace -- so that nothing but the launches follows the sync
            images = torch.empty((B, C, H, W), dtype=f32, device=dev)
            final_T = torch.empty((B, H, W), dtype=f32, device=dev)
            ncontrib = torch.empty((B, H, W), dtype=i32, device=dev)
            tr = torch.empty((B * T, 2), dtype=i32, device=dev)
            side = _side_stream(dev) if overlap else None
            grec = gfeat = cleared = None
            if need_grad and P > 0 and C > 0:
                # cleared on the side stream, under the preprocess / sort / forward blend (the 48 B per Gaussian and
                # view memset leaves the critical path)
                grec = torch.empty((B, Pp, 8), dtype=f32, device=dev)
                gfeat = torch.empty((B, Pp, cpad), dtype=f32, device=dev)
                with torch.cuda.stream(side if side is not None else main):
                    if side is not None:
                        side.wait_stream(main)
                    grec.zero_()
                    gfeat.zero_()
                    cleared = side.record_event() if side is not None else None
            guess_key = (torch.device(dev).index, P, W, H, tuple(chunks))
            guess = _m_guess().get(guess_key)
            spec = None
            if guess is not None:
                spec = [(torch.empty((mc,), dtype=i32, device=dev),
                         torch.empty((L.msb_sort_workspace_bytes_views(Pp, nb, mc, W, H),), dtype=torch.uint8, device=dev))
                        for (b0, nb), mc in zip(chunks, guess)]
            done.synchronize()
            Ms = [int(totals[b]) for b in range(B)]
            chunks0 = chunks
            chunks = _split_for_sort(chunks, Ms)  # raises before anything else is queued
            if chunks != chunks0:
                spec = None
            # next call: capacity = this call's M per chunk + 6 % (training changes M slowly)
            _m_guess()[guess_key] = [int(sum(Ms[b0:b0 + nb]) * 1.0625) + 4096 for b0, nb in chunks0]
            # phase B: one sort (side stream when several chunks overlap) + one blend grid per chunk
            two_stream = side is not None and len(chunks) > 1
            if two_stream:
                side.wait_stream(main)  # the per-view tensors were produced on `main`
            ids_all, keep = [], []
            for k, (b0, nb) in enumerate(chunks):
                M = sum(Ms[b0:b0 + nb])
                if spec is not None and M <= spec[k][0].numel():
                    ids, ws2 = spec[k][0][:M], spec[k][1]
                else:
                    ids = torch.empty((M,), dtype=i32, device=dev)
                    ws2 = torch.empty((L.msb_sort_workspace_bytes_views(Pp, nb, M, W, H),), dtype=torch.uint8,
                                      device=dev)
                nk = L.msb_sort_num_passes_views(W, H, nb) + 4 if (M > 0 and P > 0) else 0  # + keygen, offsets, duplicate, ranges
                with torch.cuda.stream(side if two_stream else main):
                    _lib.call("sort_gaussian", nk, L.msb_sort_gaussian_views, dev, ptr(uv[b0]), ptr(depth[b0]),
                              ptr(radius[b0]), ptr(tiles[b0]), Pp, nb, M, W, H, ptr(ids), ptr(tr[b0 * T:]), ptr(ws2),
                              ws2.numel(), _lib.sm_count(dev))
                    if two_stream:
                        main.wait_event(side.record_event())
                _lib.call("blend_forward", _blend_passes_fwd(cpad, C), L.msb_blend_packed_fwd_views, dev, ptr(rec[b0]),
                          ptr(featp[b0]), ptr(ids), ptr(tr[b0 * T:]), bg, C, W, H, nb, ptr(images[b0]),
                          ptr(final_T[b0]), ptr(ncontrib[b0]))
                ids_all.append(ids)
                keep.append(ws2)  # alive until both streams are joined (allocated on `main`)
            # all side-stream sorts are ordered before the last blend, hence before anything the caller enqueues
            del keep, uv, depth
            radii = radius[:, :P] if return_aux else torch.empty((0,), dtype=i32, device=dev)
            if stats is not None:
                stats["max_radii"] = radius[:, :P].amax(dim=0) if B > 0 else torch.zeros(P, dtype=i32, device=dev)
        ctx.cfg = (B, P, Pp, Cs, D, C, cpad, W, H, T, es, bg, sh_bias, clamp, with_depth)
        ctx.chunks = chunks
        ctx.cam_grad = (intrs.requires_grad, extrs.requires_grad)
        ctx.grad_sync = (grad_sync, grad_chunks)
        ctx.shapes = (tuple(opacity.shape), tuple(intrs.shape), tuple(extrs.shape))
        ctx.has_ndc = ndc is not None
        ctx.stats = stats
        ctx.gbuf = (grec, gfeat, cleared)
        ctx.gclean = True
        ctx.overlap = overlap  # backward runs on autograd's thread: the caller's serialised() does not reach it
        ctx.save_for_backward(x, s, q, sh, I, E, rec, featp, tiles, tr, final_T, ncontrib, *ids_all)
        ctx.mark_non_differentiable(radii)
        return images, radii

    @staticmethod
    def backward(ctx, dL_dimages, _dL_dradii=None):
        B, P, Pp, Cs, D, C, cpad, W, H, T, es, bg, sh_bias, clamp, with_depth = ctx.cfg
        x, s, q, sh, I, E, rec, featp, tiles, tr, final_T, ncontrib = ctx.saved_tensors[:12]
        ids_all = ctx.saved_tensors[12:]
        chunks = ctx.chunks
        g = as_f32(dL_dimages, "dL_dimages")
        dev = x.device
        L = _lib.lib()
        f32, i32 = torch.float32, torch.int32
        need_i, need_e = ctx.cam_grad
        dintr = torch.zeros((B, 4), dtype=f32, device=dev) if need_i else None
        dextr = torch.zeros((B, es), dtype=f32, device=dev) if need_e else None
        dndc = None
        op_shape = ctx.shapes[0]
        if P == 0 or C == 0:
            z = lambda *shape: torch.zeros(shape, dtype=f32, device=dev)
            if ctx.has_ndc:
                dndc = z(B, P, 2)
            return (z(P, 3), z(P, 3), z(P, 4), z(*op_shape), torch.zeros_like(sh), dintr,
                    None if dextr is None else dextr.reshape(ctx.shapes[2]), None, None, None, None, None, None, None,
                    None, None, None, dndc, None, None, None)
        group = _resolve_group(ctx.grad_sync[0])
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            grec, gfeat, cleared = ctx.gbuf
            if grec is None:
                grec = torch.empty((B, Pp, 8), dtype=f32, device=dev)
                gfeat = torch.empty((B, Pp, cpad), dtype=f32, device=dev)
                ctx.gclean = False
            if not ctx.gclean:  # a second backward through the same graph (retain_graph=True)
                grec.zero_()
                gfeat.zero_()
            elif cleared is not None:
                main.wait_event(cleared)
            ctx.gclean = False

            def blend_bwd(k):
                """backward blend of chunk k: one grid, blockIdx.z = view"""
                b0, nb = chunks[k]
                _lib.call("blend_backward", _blend_passes_bwd(cpad), L.msb_blend_packed_bwd_views, dev, ptr(rec[b0]),
                          ptr(featp[b0]), ptr(ids_all[k]), ptr(tr[b0 * T:]), bg, Pp, C, W, H, nb, ptr(final_T[b0]),
                          ptr(ncontrib[b0]), ptr(g[b0]), ptr(grec[b0]), ptr(gfeat[b0]), 1)

            def pre_bwd(v0, nv, lo, hi, accumulate, outs):
                """fused preprocess backward of the views [v0, v0 + nv) for the Gaussians [lo, hi)"""
                dxyz, dscale, dquat, dop, dshs = outs
                _lib.call("render_preprocess_backward", 1, L.msb_render_preprocess_bwd_views, dev, ptr(x[lo:hi]),
                          ptr(s[lo:hi]), ptr(q[lo:hi]), ptr(sh[lo:hi]), ptr(I[v0]), ptr(E[v0]), es,
                          ptr(tiles[v0, lo:]), ptr(grec[v0, lo:]), ptr(gfeat[v0, lo:]), hi - lo, nv, Pp, Cs, D,
                          int(with_depth), sh_bias, int(clamp), int(accumulate), ptr(dxyz[lo:hi]), ptr(dscale[lo:hi]),
                          ptr(dquat[lo:hi]), ptr(dop[lo:hi]), ptr(dshs[lo:hi]), ptr(dintr[v0]) if need_i else None,
                          ptr(dextr[v0]) if need_e else None)

            dxyz = torch.empty((P, 3), dtype=f32, device=dev)
            dscale = torch.empty((P, 3), dtype=f32, device=dev)
            dquat = torch.empty((P, 4), dtype=f32, device=dev)
            dop = torch.empty((P,), dtype=f32, device=dev)
            dshs = torch.empty_like(sh)
            outs = (dxyz, dscale, dquat, dop, dshs)
            if group is None:
                # one chunk (the default): one backward blend grid, one fused preprocess backward launch.  Several
                # chunks: the preprocess backward of chunk k (HBM-bound) on the side stream under the backward blend
                # of chunk k + 1 (issue-bound)
                side = _side_stream(dev) if (ctx.overlap and len(chunks) > 1) else None
                if side is not None:
                    side.wait_stream(main)  # the output tensors were allocated (and maybe recycled) on `main`
                for k, (b0, nb) in enumerate(chunks):
                    blend_bwd(k)
                    with torch.cuda.stream(side if side is not None else main):
                        if side is not None:
                            side.wait_event(main.record_event())
                        pre_bwd(b0, nb, 0, P, k > 0, outs)
                if side is not None:
                    main.wait_stream(side)
            else:
                # view-batch data parallelism: all backward blends, then the preprocess backward slab by
                # slab over the Gaussians; the sum all-reduce of a finished slab (its five tensors as ONE coalesced
                # NCCL call, on NCCL's stream) runs under the kernels of the next slab
                for k in range(len(chunks)):
                    blend_bwd(k)
                nslab = int(ctx.grad_sync[1]) or _default_grad_chunks(group)
                nslab = max(1, min(nslab, (P + 255) // 256))
                step = ((P + nslab - 1) // nslab + 255) // 256 * 256  # slab starts stay 16-byte aligned
                works = []
                for lo in range(0, P, step):
                    hi = min(P, lo + step)
                    for k, (b0, nb) in enumerate(chunks):
                        pre_bwd(b0, nb, lo, hi, k > 0, outs)
                    works += _reduce_many(group, [t[lo:hi] for t in outs], dev)
                for w in works:
                    if w is not None and hasattr(w, "wait"):
                        w.wait()
                # NCCL's work objects keep the reduced tensors referenced for a while; autograd's AccumulateGrad then
                # cannot adopt them as .grad and clones them (708 MB of device copies per step at 3M Gaussians).
                # Fresh aliases have a single owner and are adopted as they are.
                dxyz, dscale, dquat, dop, dshs = (t.view_as(t) for t in outs)
            gr = grec[:, :P]
            if ctx.has_ndc:  # screen-space gradient hook: dL_duv of view b is columns 0:2 of its packed record
                dndc = gr[:, :, :2] * torch.tensor([0.5 * W, 0.5 * H], dtype=f32, device=dev)
            if ctx.stats is not None:
                nd = gr[:, :, :2] * torch.tensor([0.5 * W, 0.5 * H], dtype=f32, device=dev)
                ctx.stats["ndc_grad_norm_sum"] = nd.norm(dim=-1).sum(dim=0)
                ctx.stats["ndc_grad_count"] = (tiles[:, :P] > 0).sum(dim=0).to(i32)
        return (dxyz, dscale, dquat, dop.reshape(op_shape), dshs, dintr,
                None if dextr is None else dextr.reshape(ctx.shapes[2]), None, None, None, None, None, None, None, None,
                None, None, dndc, None, None, None)


def _blend_passes_fwd(cpad: int, C: int) -> int:
    return 1 if C == 0 else (cpad // 32 + (1 if cpad % 32 else 0))


def _blend_passes_bwd(cpad: int) -> int:
    return (cpad + 15) // 16 if cpad > 8 else 1
