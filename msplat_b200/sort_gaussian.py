"""sort_gaussian: duplicate Gaussians per touched tile, sort by [tile | depth], tile ranges.

Reference: /root/reference/msplat/sort_gaussian.py:8-54 (cumsum -> compute_gaussian_key ->
torch.sort -> gather -> compute_tile_gaussian_range), src/sort_gaussian.cu:17-142 (K9/K10).
Outputs are bit-exact with the reference: idx_sorted int32 [M], tile_range int32 [T,2].
Not differentiable (like the reference).
"""
from typing import Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import as_f32, as_i32, ptr


def sort_gaussian(uv: Tensor, depth: Tensor, W: int, H: int, radius: Tensor, tiles: Tensor) -> Tuple[Tensor, Tensor]:
    """uv [P,2], depth [P,1], radius/tiles int32 [P] or [P,1] -> (idx_sorted [M], tile_range [T,2])."""
    with torch.no_grad():
        u, d = as_f32(uv, "uv"), as_f32(depth, "depth")
        r, t = as_i32(radius, "radius"), as_i32(tiles, "tiles")
        P = u.shape[0]
        if u.shape != (P, 2) or d.numel() != P or r.numel() != P or t.numel() != P:
            raise RuntimeError("uv must be [P,2]; depth, radius, tiles must have P elements")
        dev = u.device
        L = _lib.lib()
        T = ((W + 15) // 16) * ((H + 15) // 16)
        tile_range = torch.empty((T, 2), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            ws1 = torch.empty((L.msb_sort_scan_workspace_bytes(P),), dtype=torch.uint8, device=dev)
            total = _lib.pinned_i64(dev)
            _lib.call("sort_scan", 2 if P else 0, L.msb_sort_scan, dev, ptr(t), P, None, ptr(total), ptr(ws1),
                      ws1.numel())
            # the one host<->device sync of the pipeline: M sizes the output (reference: two .item())
            torch.cuda.current_stream(dev).synchronize()
            M = int(total[0])
            if M > 2 ** 31 - 1:  # int32 positions: the bound of the reference's int32 cumsum (sort_gaussian.py:42)
                raise RuntimeError(f"sort_gaussian: {M} tile intersections exceed the supported 2^31 - 1")
            idx_sorted = torch.empty((M,), dtype=torch.int32, device=dev)
            ws2 = torch.empty((L.msb_sort_workspace_bytes(P, M, int(W), int(H)),), dtype=torch.uint8, device=dev)
            nk = 4 + L.msb_sort_num_passes(int(W), int(H)) if (M > 0 and P > 0) else 0  # keygen, offsets, duplicate, ranges + passes
            _lib.call("sort_gaussian", nk, L.msb_sort_gaussian, dev, ptr(u), ptr(d), ptr(r), ptr(t), P,
                      M, int(W), int(H), ptr(idx_sorted), ptr(tile_range), ptr(ws2), ws2.numel(), _lib.sm_count(dev))
    return idx_sorted, tile_range


def sort_gaussian_views(uv: Tensor, depth: Tensor, W: int, H: int, radius: Tensor, tiles: Tensor) -> Tuple[Tensor, Tensor]:
    """Extension: ONE sort for a batch of B views of the same P Gaussians (csrc/sort.cu, keygen_kernel).
    uv [B,P,2], depth [B,P], radius/tiles int32 [B,P], P a multiple of 4 when B > 1 ->
    (idx_sorted [M] holding view * P + index, tile_range [B,T,2] indexing idx_sorted), M = tiles.sum().
    Within a view the order is exactly that of :func:`sort_gaussian`."""
    with torch.no_grad():
        u, d = as_f32(uv, "uv"), as_f32(depth, "depth")
        r, t = as_i32(radius, "radius"), as_i32(tiles, "tiles")
        if u.dim() != 3 or u.shape[2] != 2:
            raise RuntimeError("uv must be [B,P,2]")
        B, P = int(u.shape[0]), int(u.shape[1])
        if d.numel() != B * P or r.numel() != B * P or t.numel() != B * P:
            raise RuntimeError("depth, radius, tiles must have B*P elements")
        if B > 1 and P % 4:
            raise RuntimeError("sort_gaussian_views: P must be a multiple of 4 (pad with tiles = 0)")
        dev = u.device
        L = _lib.lib()
        T = ((W + 15) // 16) * ((H + 15) // 16)
        tile_range = torch.empty((B, T, 2), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            ws1 = torch.empty((L.msb_sort_scan_workspace_bytes(B * P),), dtype=torch.uint8, device=dev)
            total = _lib.pinned_i64(dev)
            _lib.call("sort_scan", 2 if B * P else 0, L.msb_sort_scan, dev, ptr(t), B * P, None, ptr(total), ptr(ws1),
                      ws1.numel())
            torch.cuda.current_stream(dev).synchronize()
            M = int(total[0])
            if M > 2 ** 31 - 1:
                raise RuntimeError(f"sort_gaussian_views: {M} tile intersections exceed the supported 2^31 - 1")
            idx_sorted = torch.empty((M,), dtype=torch.int32, device=dev)
            ws2 = torch.empty((L.msb_sort_workspace_bytes_views(P, B, M, int(W), int(H)),), dtype=torch.uint8, device=dev)
            nk = L.msb_sort_num_passes_views(int(W), int(H), B) + 4 if (M > 0 and P > 0) else 0
            _lib.call("sort_gaussian", nk, L.msb_sort_gaussian_views, dev, ptr(u), ptr(d), ptr(r), ptr(t), P, B,
                      M, int(W), int(H), ptr(idx_sorted), ptr(tile_range), ptr(ws2), ws2.numel(), _lib.sm_count(dev))
    return idx_sorted, tile_range
