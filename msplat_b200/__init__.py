"""msplat_b200 -- a B200-native (sm_100a) differentiable Gaussian-splatting rasterizer that is a
drop-in for the Python API of pointrix-project/msplat.

Same seven public names and signatures as /root/reference/msplat/__init__.py:11-19:
``rasterization`` plus the steps interface ``project_point``, ``compute_cov3d``, ``ewa_project``,
``compute_sh``, ``sort_gaussian``, ``alpha_blending``.  Every step is a ``torch.autograd.Function``
over hand-written CUDA kernels reached through the C ABI in ``include/msplat_b200.h``.
"""
import torch
from torch import Tensor

from . import _lib, optim  # noqa: F401  (msplat_b200.optim.FusedAdam)
from ._lib import as_f32, ptr
from .alpha_blending import _blend_backward, _blend_forward, alpha_blending
from .compute_cov3d import compute_cov3d
from .compute_sh import compute_sh
from .ewa_project import ewa_project
from .project_point import project_point
from .render import rasterization_sh, rasterization_sh_views
from .sort_gaussian import sort_gaussian, sort_gaussian_views

__all__ = [
    "project_point",
    "compute_cov3d",
    "ewa_project",
    "sort_gaussian",
    "compute_sh",
    "alpha_blending",
    "rasterization",
    # extensions beyond the reference API (fused SH render path, csrc/render.cu)
    "rasterization_sh",
    "rasterization_sh_views",
    "sort_gaussian_views",
]

__version__ = "0.1.0"


def rasterization(
    xyz: Tensor, scale: Tensor, rotate: Tensor, opacity: Tensor, feature: Tensor, intr: Tensor, extr: Tensor,
    W: int, H: int, bg: float, ndc: Tensor = None, *, fused: bool = True,
) -> Tensor:
    """Vanilla 3D Gaussian Splatting rasterization pipeline -> feature map [C, H, W].

    Mirrors /root/reference/msplat/__init__.py:22-93: project -> ``visible = depth != 0`` ->
    cov3d -> ewa -> sort -> blend.  With ``fused=True`` (default) the per-Gaussian stages run as
    one forward and one backward kernel inside a single autograd Function (no cov3d tensor, no
    intermediate autograd nodes); results are bit-identical to the step-by-step composition
    (``fused=False``), which is kept for parity testing.
    """
    if fused:
        return _Rasterize.apply(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg, ndc)
    (uv, depth) = project_point(xyz, intr, extr, W, H)
    visible = depth != 0
    cov3d = compute_cov3d(scale, rotate, visible)
    (conic, radius, tiles_touched) = ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
    (gaussian_ids_sorted, tile_range) = sort_gaussian(uv, depth, W, H, radius, tiles_touched)
    return alpha_blending(uv, conic, opacity, feature, gaussian_ids_sorted, tile_range, bg, W, H, ndc)


class _Rasterize(torch.autograd.Function):
    """Fused pipeline."""

    @staticmethod
    def forward(ctx, xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg, ndc):
        x, s, q = as_f32(xyz, "xyz"), as_f32(scale, "scale"), as_f32(rotate, "rotate")
        o, f = as_f32(opacity, "opacity"), as_f32(feature, "feature")
        i, e = as_f32(intr, "intr"), as_f32(extr, "extr")
        P = x.shape[0]
        if x.shape != (P, 3) or s.shape != (P, 3) or q.shape != (P, 4) or o.numel() != P or f.dim() != 2 \
                or f.shape[0] != P:
            raise RuntimeError("rasterization: xyz [P,3], scale [P,3], rotate [P,4], opacity [P,1], feature [P,C]")
        dev = x.device
        uv = torch.empty((P, 2), dtype=torch.float32, device=dev)
        depth = torch.empty((P, 1), dtype=torch.float32, device=dev)
        conic = torch.empty((P, 3), dtype=torch.float32, device=dev)
        radius = torch.empty((P,), dtype=torch.int32, device=dev)
        tiles = torch.empty((P,), dtype=torch.int32, device=dev)
        _lib.call("preprocess_forward", 1 if P else 0, _lib.lib().msb_preprocess_fwd, dev, ptr(x), ptr(s), ptr(q),
                  ptr(i), ptr(e), P, int(W), int(H), 0.0, 1.3, ptr(uv), ptr(depth), ptr(conic), ptr(radius), ptr(tiles))
        ids, tr = sort_gaussian(uv, depth, W, H, radius, tiles)
        image, final_T, ncontrib, packed = _blend_forward(uv, conic, o, f, ids, tr, bg, W, H)
        ctx.W, ctx.H, ctx.bg = W, H, bg
        ctx.has_ndc = ndc is not None
        ctx.cam_grad = (intr.requires_grad, extr.requires_grad)
        ctx.extr_shape = tuple(extr.shape)
        ctx.opacity_shape = tuple(opacity.shape)
        ctx.save_for_backward(x, s, q, i, e, depth, radius, f, ids, tr, final_T, ncontrib, packed)
        return image

    @staticmethod
    def backward(ctx, dL_dimage):
        x, s, q, i, e, depth, radius, f, ids, tr, final_T, ncontrib, packed = ctx.saved_tensors
        W, H = ctx.W, ctx.H
        P = x.shape[0]
        dev = x.device
        g = as_f32(dL_dimage, "dL_dimage")
        dL_duv, dL_dconic, dL_dopacity, dL_dfeature = _blend_backward(f, ids, tr, ctx.bg, W, H, final_T, ncontrib, g,
                                                                     packed)
        dL_dxyz = torch.empty((P, 3), dtype=torch.float32, device=dev)
        dL_dscale = torch.empty((P, 3), dtype=torch.float32, device=dev)
        dL_dquat = torch.empty((P, 4), dtype=torch.float32, device=dev)
        need_i, need_e = ctx.cam_grad
        dL_dintr = torch.zeros(4, dtype=torch.float32, device=dev) if need_i else None
        dL_dextr = torch.zeros(ctx.extr_shape, dtype=torch.float32, device=dev) if need_e else None
        _lib.call("preprocess_backward", 1 if P else 0, _lib.lib().msb_preprocess_bwd, dev, ptr(x), ptr(s), ptr(q),
                  ptr(i), ptr(e), ptr(depth), ptr(radius), ptr(dL_duv), None, ptr(dL_dconic), P, ptr(dL_dxyz),
                  ptr(dL_dscale), ptr(dL_dquat), ptr(dL_dintr), ptr(dL_dextr))
        dL_dndc = None
        if ctx.has_ndc:
            dL_dndc = dL_duv * torch.tensor([0.5 * W, 0.5 * H], dtype=dL_duv.dtype, device=dev)[None, :]
        return (dL_dxyz, dL_dscale, dL_dquat, dL_dopacity.reshape(ctx.opacity_shape), dL_dfeature, dL_dintr, dL_dextr,
                None, None, None, dL_dndc)
