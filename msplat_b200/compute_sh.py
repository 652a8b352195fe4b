"""compute_sh: spherical harmonics (degree 0..10) -> per-Gaussian channel values.

Reference: /root/reference/msplat/compute_sh.py:8-64, src/compute_sh.cu:116-1754 (K7/K8).
shs is [P, Cs, D] with D = (deg+1)^2 innermost; no +0.5, no clamp.
"""
import torch
from torch import Tensor

from . import _lib
from ._lib import as_f32, as_mask, ptr


def compute_sh(shs: Tensor, view_dirs: Tensor, visible: Tensor = None) -> Tensor:
    """shs [P,Cs,D], view_dirs [P,3], visible [P]/[P,1] bool -> value [P,Cs]."""
    if visible is None:
        visible = torch.ones_like(shs[:, 0, 0], dtype=torch.bool)
    return _ComputeSH.apply(shs, view_dirs, visible)


class _ComputeSH(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shs, view_dirs, visible):
        s, d = as_f32(shs, "shs"), as_f32(view_dirs, "view_dirs")
        if s.dim() != 3 or d.shape != (s.shape[0], 3):
            raise RuntimeError("shs must be [P, Cs, D] and view_dirs [P, 3]")
        P, Cs, D = s.shape
        deg = int(round(D ** 0.5)) - 1
        if (deg + 1) ** 2 != D or not 0 <= deg <= 10:
            raise RuntimeError(f"shs last dim must be (deg+1)^2 with deg <= 10, got {D}")
        vis = as_mask(visible, "visible", P)
        value = torch.empty((P, Cs), dtype=torch.float32, device=s.device)
        _lib.call("compute_sh_forward", 1 if P and Cs else 0, _lib.lib().msb_compute_sh_fwd, s.device, ptr(s), ptr(d),
                  ptr(vis), P, Cs, D, ptr(value))
        ctx.save_for_backward(s, d, vis)
        return value

    @staticmethod
    def backward(ctx, dL_dvalue):
        s, d, vis = ctx.saved_tensors
        P, Cs, D = s.shape
        g = as_f32(dL_dvalue, "dL_dvalue")
        dL_dshs = torch.empty_like(s)
        dL_ddirs = torch.empty((P, 3), dtype=torch.float32, device=s.device)
        if Cs == 0:
            dL_ddirs.zero_()
        else:
            _lib.call("compute_sh_backward", 1 if P else 0, _lib.lib().msb_compute_sh_bwd, s.device, ptr(s), ptr(d),
                      ptr(vis), ptr(g), P, Cs, D, ptr(dL_dshs), ptr(dL_ddirs))
        return dL_dshs, dL_ddirs, None
