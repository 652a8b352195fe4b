"""alpha_blending: tile-based front-to-back blending of arbitrary-channel features.

Reference: /root/reference/msplat/alpha_blending.py:7-135, src/alpha_blending.cu:16-573
(K11/K12 + the channel-chunk dispatcher D1).
"""
import torch
from torch import Tensor

from . import _lib
from ._lib import as_f32, as_i32, ptr


def alpha_blending(
    uv: Tensor, conic: Tensor, opacity: Tensor, feature: Tensor, idx_sorted: Tensor, tile_range: Tensor,
    bg: float, W: int, H: int, ndc: Tensor = None,
) -> Tensor:
    """uv [P,2], conic [P,3], opacity [P,1], feature [P,C], idx_sorted int32 [M], tile_range int32
    [T,2] -> feature map [C,H,W].  ``ndc`` [P,2] only receives dL_duv * (0.5 W, 0.5 H)."""
    return _AlphaBlending.apply(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, ndc)


def _blend_forward(u, c, o, f, ids, tr, bg, W, H):
    P, C = f.shape
    dev = f.device
    L = _lib.lib()
    image = torch.empty((C, H, W), dtype=torch.float32, device=dev)
    final_T = torch.empty((H, W), dtype=torch.float32, device=dev)
    ncontrib = torch.empty((H, W), dtype=torch.int32, device=dev)
    packed = torch.empty((L.msb_blend_fwd_workspace_bytes(P, C),), dtype=torch.uint8, device=dev)
    cpad = L.msb_blend_cpad(C)
    npass = 1 if C == 0 else (cpad // 32 + (1 if cpad % 32 else 0))
    _lib.call("alpha_blending_forward", (1 if P else 0) + npass, L.msb_alpha_blending_fwd, dev, ptr(u), ptr(c), ptr(o),
              ptr(f), ptr(ids), ptr(tr), float(bg), P, C, int(W), int(H), ptr(image), ptr(final_T), ptr(ncontrib),
              ptr(packed), packed.numel())
    return image, final_T, ncontrib, packed


def _blend_backward(f, ids, tr, bg, W, H, final_T, ncontrib, g, packed):
    P, C = f.shape
    dev = f.device
    L = _lib.lib()
    dL_duv = torch.empty((P, 2), dtype=torch.float32, device=dev)
    dL_dconic = torch.empty((P, 3), dtype=torch.float32, device=dev)
    dL_dopacity = torch.empty((P, 1), dtype=torch.float32, device=dev)
    dL_dfeature = torch.empty((P, C), dtype=torch.float32, device=dev)
    if P == 0 or C == 0:
        for t in (dL_duv, dL_dconic, dL_dopacity, dL_dfeature):
            t.zero_()
        return dL_duv, dL_dconic, dL_dopacity, dL_dfeature
    ws = torch.empty((L.msb_blend_bwd_workspace_bytes(P, C),), dtype=torch.uint8, device=dev)
    cpad = L.msb_blend_cpad(C)
    nk = 1 + ((cpad + 15) // 16 if cpad > 8 else 1)
    _lib.call("alpha_blending_backward", nk, L.msb_alpha_blending_bwd, dev, ptr(f), ptr(ids), ptr(tr), float(bg), P, C,
              int(W), int(H), ptr(final_T), ptr(ncontrib), ptr(g), ptr(packed), ptr(dL_duv), ptr(dL_dconic),
              ptr(dL_dopacity), ptr(dL_dfeature), ptr(ws), ws.numel())
    return dL_duv, dL_dconic, dL_dopacity, dL_dfeature


class _AlphaBlending(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, ndc):
        u, c, o, f = as_f32(uv, "uv"), as_f32(conic, "conic"), as_f32(opacity, "opacity"), as_f32(feature, "feature")
        ids, tr = as_i32(idx_sorted, "idx_sorted"), as_i32(tile_range, "tile_range")
        if f.dim() != 2:
            raise RuntimeError("feature must be [P, C]")
        P = f.shape[0]
        T = ((W + 15) // 16) * ((H + 15) // 16)
        if u.shape != (P, 2) or c.shape != (P, 3) or o.numel() != P or tr.numel() != 2 * T:
            raise RuntimeError("alpha_blending: uv [P,2], conic [P,3], opacity [P,1], tile_range [T,2] expected")
        image, final_T, ncontrib, packed = _blend_forward(u, c, o, f, ids, tr, bg, W, H)
        ctx.W, ctx.H, ctx.bg = W, H, bg
        ctx.has_ndc = ndc is not None
        ctx.opacity_shape = tuple(opacity.shape)
        ctx.save_for_backward(f, ids, tr, final_T, ncontrib, packed)
        return image

    @staticmethod
    def backward(ctx, dL_drendered):
        f, ids, tr, final_T, ncontrib, packed = ctx.saved_tensors
        W, H = ctx.W, ctx.H
        g = as_f32(dL_drendered, "dL_drendered")
        dL_duv, dL_dconic, dL_dopacity, dL_dfeature = _blend_backward(f, ids, tr, ctx.bg, W, H, final_T, ncontrib, g,
                                                                     packed)
        dL_dndc = None
        if ctx.has_ndc:  # alpha_blending.py:107-110 of the reference
            dL_dndc = dL_duv * torch.tensor([0.5 * W, 0.5 * H], dtype=dL_duv.dtype, device=dL_duv.device)[None, :]
        return (dL_duv, dL_dconic, dL_dopacity.reshape(ctx.opacity_shape), dL_dfeature, None, None, None, None, None,
                dL_dndc)
