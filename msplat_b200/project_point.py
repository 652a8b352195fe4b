"""project_point: 3-D points -> pixel coordinates + depth (drop-in for msplat.project_point).

Reference: /root/reference/msplat/project_point.py:8-98 (API, autograd wrapper),
/root/reference/msplat/src/project_point.cu:13-145 (kernels K1/K2).
"""
from typing import Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import as_f32, ptr


def project_point(
    xyz: Tensor, intr: Tensor, extr: Tensor, W: int, H: int, nearest: float = 0.0, extent: float = 1.3
) -> Tuple[Tensor, Tensor]:
    """Project 3D points to the screen.

    xyz [P,3]; intr [4] = (fx, fy, cx, cy); extr [3,4] (or [4,4]: only the first 12 floats are
    read); returns uv [P,2] and depth [P,1].  Culled points (near plane only if ``nearest > 0``,
    image extent ``extent``) have uv = depth = 0.
    """
    return _ProjectPoint.apply(xyz, intr, extr, W, H, nearest, extent)


class _ProjectPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, intr, extr, W, H, nearest, extent):
        xyz_c, intr_c, extr_c = as_f32(xyz, "xyz"), as_f32(intr, "intr"), as_f32(extr, "extr")
        if xyz_c.dim() != 2 or xyz_c.shape[1] != 3:
            raise RuntimeError(f"xyz must be [P, 3], got {tuple(xyz.shape)}")
        if intr_c.numel() < 4 or extr_c.numel() < 12:
            raise RuntimeError("intr must have 4 elements and extr at least 12")
        P = xyz_c.shape[0]
        uv = torch.empty((P, 2), dtype=torch.float32, device=xyz_c.device)
        depth = torch.empty((P, 1), dtype=torch.float32, device=xyz_c.device)
        _lib.call("project_point_forward", 1 if P else 0, _lib.lib().msb_project_point_fwd, xyz_c.device,
                  ptr(xyz_c), ptr(intr_c), ptr(extr_c), P, int(W), int(H), float(nearest), float(extent), ptr(uv),
                  ptr(depth))
        ctx.W, ctx.H = W, H
        ctx.cam_grad = (intr.requires_grad, extr.requires_grad)
        ctx.extr_shape = tuple(extr.shape)
        ctx.save_for_backward(xyz_c, intr_c, extr_c, depth)
        return uv, depth

    @staticmethod
    def backward(ctx, dL_duv, dL_ddepth):
        xyz, intr, extr, depth = ctx.saved_tensors
        P = xyz.shape[0]
        dev = xyz.device
        g_uv, g_d = as_f32(dL_duv, "dL_duv"), as_f32(dL_ddepth, "dL_ddepth")
        dL_dxyz = torch.empty((P, 3), dtype=torch.float32, device=dev)
        need_i, need_e = ctx.cam_grad
        dL_dintr = torch.zeros(4, dtype=torch.float32, device=dev) if need_i else None
        dL_dextr = torch.zeros(ctx.extr_shape, dtype=torch.float32, device=dev) if need_e else None
        _lib.call("project_point_backward", 1 if P else 0, _lib.lib().msb_project_point_bwd, dev, ptr(xyz),
                  ptr(intr), ptr(extr), ptr(depth), ptr(g_uv), ptr(g_d), P, ptr(dL_dxyz), ptr(dL_dintr), ptr(dL_dextr))
        return dL_dxyz, dL_dintr, dL_dextr, None, None, None, None
