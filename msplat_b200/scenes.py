"""Deterministic synthetic Gaussian clouds for tests and bench.py.

All scenes are generated on the CPU with a seeded ``torch.Generator`` and then moved, so the
same tensors can be fed to this library, to the oracle and to the reference build.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class Scene:
    xyz: torch.Tensor      # [P,3]
    scale: torch.Tensor    # [P,3]
    quat: torch.Tensor     # [P,4] (r,x,y,z), normalised
    opacity: torch.Tensor  # [P,1]
    shs: torch.Tensor      # [P,Cs,D] or None
    intr: torch.Tensor     # [4]
    extr: torch.Tensor     # [3,4]
    W: int
    H: int

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)
        return Scene(mv(self.xyz), mv(self.scale), mv(self.quat), mv(self.opacity), mv(self.shs), mv(self.intr),
                     mv(self.extr), self.W, self.H)

    @property
    def cam_center(self):
        R, t = self.extr[:3, :3], self.extr[:3, 3]
        return -(R.T @ t)


def frustum_scene(P: int, W: int, H: int, sigma_med: float = 2.0, seed: int = 0, sh_degree: int = 3,
                  sh_channels: int = 3, with_sh: bool = True) -> Scene:
    """S-frustum(P, W, H, sigma_med, seed): fov_x 60 deg, extr = [I|0]; pixel positions uniform
    over the image plus a 10 % margin, z log-uniform in [1, 50] (all z > 0); pixel
    footprint sigma_px = sigma_med * exp(0.8 N(0,1)); anisotropic scales; random unit quaternions;
    opacity U(0,1); SH DC ~ 0.5 N, higher orders ~ 0.1 N."""
    g = torch.Generator().manual_seed(seed)
    fx = fy = 0.5 * W / math.tan(math.radians(30.0))
    cx, cy = W / 2.0, H / 2.0
    u = (torch.rand(P, generator=g) * 1.2 - 0.1) * W
    v = (torch.rand(P, generator=g) * 1.2 - 0.1) * H
    z = torch.exp(torch.rand(P, generator=g) * math.log(50.0))
    xyz = torch.stack([(u + 0.5 - cx) * z / fx, (v + 0.5 - cy) * z / fy, z], dim=-1)
    sigma_px = sigma_med * torch.exp(0.8 * torch.randn(P, generator=g))
    scale = (sigma_px * z / fx)[:, None] * (0.3 + 0.7 * torch.rand(P, 3, generator=g))
    quat = torch.randn(P, 4, generator=g)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    opacity = torch.rand(P, 1, generator=g)
    shs = None
    if with_sh:
        D = (sh_degree + 1) ** 2
        shs = 0.1 * torch.randn(P, sh_channels, D, generator=g)
        shs[:, :, 0] = 0.5 * torch.randn(P, sh_channels, generator=g)
    intr = torch.tensor([fx, fy, cx, cy], dtype=torch.float32)
    extr = torch.cat([torch.eye(3), torch.zeros(3, 1)], dim=1)
    return Scene(xyz.float(), scale.float(), quat.float(), opacity.float(), shs, intr, extr, W, H)


def cube_scene(P: int = 10000, W: int = 256, H: int = 256, seed: int = 0, sh_degree: int = 3,
               scale_lo: float = 0.005, scale_hi: float = 0.05) -> Scene:
    """BASELINE config #1 (gs_2d-style): xyz ~ U(-1,1)^3, extr = [I | (0,0,2.5)], fov 90 deg."""
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(P, 3, generator=g) * 2 - 1
    scale = scale_lo + (scale_hi - scale_lo) * torch.rand(P, 3, generator=g)
    quat = torch.randn(P, 4, generator=g)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    opacity = torch.rand(P, 1, generator=g)
    D = (sh_degree + 1) ** 2
    shs = 0.1 * torch.randn(P, 3, D, generator=g)
    shs[:, :, 0] = 0.5 * torch.randn(P, 3, generator=g)
    f = 0.5 * W / math.tan(math.radians(45.0))
    intr = torch.tensor([f, f, W / 2.0, H / 2.0], dtype=torch.float32)
    extr = torch.cat([torch.eye(3), torch.tensor([[0.0], [0.0], [2.5]])], dim=1)
    return Scene(xyz, scale, quat, opacity, shs, intr, extr, W, H)


def bunny2d_scene(P: int = 100000, W: int = 512, H: int = 512, seed: int = 123) -> Scene:
    """BASELINE config #2: the initialisation of /root/reference/tutorials/gs_2d.py:10-27,57-64
    (fov 90 deg, t = (0,0,2.5), scale = |U(0,1)| + 1e-8, rotate = normalize(U(0,1)^4),
    opacity = sigmoid(U(0,1)); colours are passed as features by the caller)."""
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(P, 3, generator=g) * 2 - 1)
    scale = torch.rand(P, 3, generator=g).abs() + 1e-8
    quat = torch.rand(P, 4, generator=g)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    opacity = torch.sigmoid(torch.rand(P, 1, generator=g))
    f = 0.5 * W / math.tan(math.radians(45.0))
    intr = torch.tensor([f, f, W / 2.0, H / 2.0], dtype=torch.float32)
    extr = torch.cat([torch.eye(3), torch.tensor([[0.0], [0.0], [2.5]])], dim=1)
    return Scene(xyz, scale, quat, opacity, None, intr, extr, W, H)


def orbit_cameras(n: int = 64, yaw_deg: float = 20.0, shift: float = 1.0):
    """BASELINE config #5 cameras: yaw (k/(n-1) - 0.5) * yaw_deg about y plus an x-translation of
    (k/(n-1) - 0.5) * shift.  Returns a list of [3,4] extrinsics."""
    out = []
    for k in range(n):
        a = (k / max(n - 1, 1) - 0.5)
        th = math.radians(a * yaw_deg)
        R = torch.tensor([[math.cos(th), 0.0, math.sin(th)], [0.0, 1.0, 0.0], [-math.sin(th), 0.0, math.cos(th)]])
        t = torch.tensor([[a * shift], [0.0], [0.0]])
        out.append(torch.cat([R, t], dim=1).float())
    return out
