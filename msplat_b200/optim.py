"""Fused multi-tensor Adam (csrc/adam.cu): the optimizer step of the reference's training loop
(/root/reference/tutorials/gs_2d.py:32-36: ``torch.optim.Adam(list(attributes), lr=0.01)``; :40-42 ``step()`` +
``zero_grad()``) as ONE kernel launch over all parameter tensors instead of torch's per-tensor elementwise kernels.
Same update rule and defaults as ``torch.optim.Adam`` (betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).
"""
from __future__ import annotations

import ctypes
from typing import Iterable

import torch

from . import _lib

__all__ = ["FusedAdam"]


class FusedAdam:
    """Drop-in for ``torch.optim.Adam(params, lr)`` on float32 CUDA parameters: ``step()`` / ``zero_grad()``.
    Parameters without a gradient are skipped (like torch)."""

    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params = list(params)
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam: parameters must be contiguous float32 CUDA tensors")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.steps = 0

    @torch.no_grad()
    def step(self):
        idx = [k for k, p in enumerate(self.params) if p.grad is not None]
        if not idx:
            return
        self.steps += 1
        n = len(idx)
        arr = lambda vals: (ctypes.c_void_p * n)(*vals)
        grads = []
        for k in idx:
            g = self.params[k].grad
            if g.dtype != torch.float32 or not g.is_cuda:
                raise RuntimeError("FusedAdam: gradients must be float32 CUDA tensors")
            grads.append(g.contiguous())
        dev = self.params[idx[0]].device
        _lib.call("adam_step", (n + 7) // 8, _lib.lib().msb_adam_step, dev, n,
                  arr([self.params[k].data_ptr() for k in idx]), arr([g.data_ptr() for g in grads]),
                  arr([self.exp_avg[k].data_ptr() for k in idx]), arr([self.exp_avg_sq[k].data_ptr() for k in idx]),
                  (ctypes.c_longlong * n)(*[self.params[k].numel() for k in idx]), self.lr, self.betas[0],
                  self.betas[1], self.eps, self.steps)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()
