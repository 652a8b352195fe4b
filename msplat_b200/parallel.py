"""View-batch data parallelism: the reference has no distributed code; the path
shards naturally across camera views.  Every rank holds a full replica of the Gaussians, renders
its contiguous slice of the view batch, accumulates per-Gaussian gradients locally into ONE flat
buffer, and a single sum all-reduce (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests)
makes the gradients identical on all ranks.  There is no other collective on the data path.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> range:
    """Contiguous slice of the view batch owned by `rank` (remainder spread over the first ranks)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n_views, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class FlatGrads:
    """One contiguous gradient buffer whose slices are the ``.grad`` of the given leaf tensors, so
    that local accumulation over views happens in place and the step needs ONE all-reduce.

    The ``.grad`` views must stay attached: clear gradients with :meth:`zero_` (or
    ``optimizer.zero_grad(set_to_none=False)``), not with ``zero_grad()``'s default ``set_to_none=True``, which
    would detach them -- :meth:`all_reduce` checks this and raises instead of silently reducing stale zeros."""

    def __init__(self, params: Sequence[torch.Tensor]):
        self.params = list(params)
        if not self.params:
            raise ValueError("no parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        sizes = [p.numel() for p in self.params]
        # keep every slice 16-byte aligned
        self.offsets: List[Tuple[int, int]] = []
        off = 0
        for n in sizes:
            self.offsets.append((off, n))
            off += (n + 3) // 4 * 4
        self.flat = torch.zeros(off, dtype=dt, device=dev)
        self.attach()

    def attach(self):
        for p, (o, n) in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + n].view_as(p)

    def zero_(self):
        self.flat.zero_()

    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def attached(self) -> bool:
        """every parameter's .grad is still its slice of the flat buffer"""
        base, esz = self.flat.data_ptr(), self.flat.element_size()
        return all(p.grad is not None and p.grad.data_ptr() == base + o * esz for p, (o, n) in zip(self.params, self.offsets))

    def all_reduce(self, async_op: bool = False):
        if not self.attached():
            raise RuntimeError("FlatGrads: a parameter's .grad no longer points into the flat buffer (zero_grad("
                               "set_to_none=True)?): call attach() before the backward passes, clear with zero_()")
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return None


def render_view_batch(render_one, cameras: Sequence, rank: int, world: int, grads: FlatGrads):
    """Render this rank's views with ``render_one(camera) -> scalar loss`` (which must call
    backward itself or return a loss to back-propagate), then all-reduce the flat gradients.
    Returns the summed local loss (a tensor)."""
    total = None
    for k in shard_views(len(cameras), rank, world):
        loss = render_one(cameras[k])
        if loss.requires_grad:
            loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    grads.all_reduce()
    return total
