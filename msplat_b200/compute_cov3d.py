"""compute_cov3d: scale + quaternion -> upper triangle of the 3-D covariance.

Reference: /root/reference/msplat/compute_cov3d.py:7-64, src/compute_cov3d.cu:14-147 (K3/K4).
Quaternion order is (r, x, y, z) and is NOT normalised inside.
"""
import torch
from torch import Tensor

from . import _lib
from ._lib import as_f32, as_mask, ptr


def compute_cov3d(scales: Tensor, uquats: Tensor, visible: Tensor = None) -> Tensor:
    """scales [P,3], uquats [P,4], visible [P] or [P,1] bool (default: all) -> cov3d [P,6]."""
    if visible is None:
        visible = torch.ones_like(scales[:, 0], dtype=torch.bool)
    return _ComputeCov3D.apply(scales, uquats, visible)


class _ComputeCov3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scales, uquats, visible):
        s, q = as_f32(scales, "scales"), as_f32(uquats, "uquats")
        if s.dim() != 2 or s.shape[1] != 3 or q.shape != (s.shape[0], 4):
            raise RuntimeError("scales must be [P, 3] and uquats [P, 4]")
        P = s.shape[0]
        vis = as_mask(visible, "visible", P)
        cov3d = torch.empty((P, 6), dtype=torch.float32, device=s.device)
        _lib.call("compute_cov3d_forward", 1 if P else 0, _lib.lib().msb_compute_cov3d_fwd, s.device, ptr(s), ptr(q),
                  ptr(vis), P, ptr(cov3d))
        ctx.save_for_backward(s, q, vis)
        return cov3d

    @staticmethod
    def backward(ctx, dL_dcov3d):
        s, q, vis = ctx.saved_tensors
        P = s.shape[0]
        g = as_f32(dL_dcov3d, "dL_dcov3d")
        dL_ds = torch.empty((P, 3), dtype=torch.float32, device=s.device)
        dL_dq = torch.empty((P, 4), dtype=torch.float32, device=s.device)
        _lib.call("compute_cov3d_backward", 1 if P else 0, _lib.lib().msb_compute_cov3d_bwd, s.device, ptr(s), ptr(q),
                  ptr(vis), ptr(g), P, ptr(dL_ds), ptr(dL_dq))
        return dL_ds, dL_dq, None
