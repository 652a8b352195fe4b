"""ewa_project: 3-D Gaussians -> 2-D conic, integer radius, tiles touched.

Reference: /root/reference/msplat/ewa_project.py:8-94, src/ewa_project.cu:16-252 (K5/K6),
include/utils.h:17-37 (get_rect).  radius and tiles are bit-exact with the reference build.
"""
from typing import Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import as_f32, as_mask, ptr


def ewa_project(
    xyz: Tensor, cov3d: Tensor, intr: Tensor, extr: Tensor, uv: Tensor, W: int, H: int, visible: Tensor = None
) -> Tuple[Tensor, Tensor, Tensor]:
    """Returns conic [P,3] float32, radius [P] int32, tiles [P] int32 (zeros where skipped)."""
    if visible is None:
        visible = torch.ones_like(uv[:, 0], dtype=torch.bool)
    return _EWAProject.apply(xyz, cov3d, intr, extr, uv, W, H, visible)


class _EWAProject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, cov3d, intr, extr, uv, W, H, visible):
        x, c = as_f32(xyz, "xyz"), as_f32(cov3d, "cov3d")
        i, e, u = as_f32(intr, "intr"), as_f32(extr, "extr"), as_f32(uv, "uv")
        P = x.shape[0]
        if x.shape != (P, 3) or c.shape != (P, 6) or u.shape != (P, 2):
            raise RuntimeError("xyz must be [P,3], cov3d [P,6], uv [P,2]")
        vis = as_mask(visible, "visible", P)
        dev = x.device
        conic = torch.empty((P, 3), dtype=torch.float32, device=dev)
        radius = torch.empty((P,), dtype=torch.int32, device=dev)
        tiles = torch.empty((P,), dtype=torch.int32, device=dev)
        _lib.call("ewa_project_forward", 1 if P else 0, _lib.lib().msb_ewa_project_fwd, dev, ptr(x), ptr(c), ptr(i),
                  ptr(e), ptr(u), ptr(vis), P, int(W), int(H), ptr(conic), ptr(radius), ptr(tiles))
        ctx.cam_grad = (intr.requires_grad, extr.requires_grad)
        ctx.extr_shape = tuple(extr.shape)
        ctx.save_for_backward(x, c, i, e, radius)
        ctx.mark_non_differentiable(radius, tiles)
        return conic, radius, tiles

    @staticmethod
    def backward(ctx, dL_dconic, dL_dradius, dL_dtiles):
        x, c, i, e, radius = ctx.saved_tensors
        P = x.shape[0]
        dev = x.device
        g = as_f32(dL_dconic, "dL_dconic")
        dL_dxyz = torch.empty((P, 3), dtype=torch.float32, device=dev)
        dL_dcov3d = torch.empty((P, 6), dtype=torch.float32, device=dev)
        need_i, need_e = ctx.cam_grad
        dL_dintr = torch.zeros(4, dtype=torch.float32, device=dev) if need_i else None
        dL_dextr = torch.zeros(ctx.extr_shape, dtype=torch.float32, device=dev) if need_e else None
        _lib.call("ewa_project_backward", 1 if P else 0, _lib.lib().msb_ewa_project_bwd, dev, ptr(x), ptr(c), ptr(i),
                  ptr(e), ptr(radius), ptr(g), P, ptr(dL_dxyz), ptr(dL_dcov3d), ptr(dL_dintr), ptr(dL_dextr))
        return dL_dxyz, dL_dcov3d, dL_dintr, dL_dextr, None, None, None, None
