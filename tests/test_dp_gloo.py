"""World-size-2 gloo test (CPU) of the view-batch data-parallel plumbing used by bench.py --gpus N:
view sharding, in-place accumulation into the flat gradient buffer, one sum all-reduce.
The per-view 'render' here is the CPU oracle's differentiable projection (no CUDA needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from msplat_b200.parallel import FlatGrads, render_view_batch, shard_views
from msplat_b200.scenes import orbit_cameras


def test_shard_views_partition():
    for n in (1, 7, 64):
        for world in (1, 2, 3, 8):
            got = [k for r in range(world) for k in shard_views(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_views(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _make_params():
    g = torch.Generator().manual_seed(0)
    xyz = (torch.randn(200, 3, generator=g) + torch.tensor([0.0, 0.0, 5.0])).requires_grad_()
    w = torch.randn(200, 2, generator=g).requires_grad_()
    return xyz, w


def _loss(xyz, w, extr):
    intr = torch.tensor([100.0, 100.0, 32.0, 32.0])
    uv, depth = oracle.project_point(xyz, intr, extr, 64, 64)
    return (uv * w).sum() * 1e-2 + depth.sum()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        xyz, w = _make_params()
        grads = FlatGrads([xyz, w])
        cams = orbit_cameras(6)
        total = render_view_batch(lambda e: _loss(xyz, w, e), cams, rank, world, grads)
        assert xyz.grad.data_ptr() == grads.flat.data_ptr()  # grads accumulated in place in the flat buffer
        # slab count of the pipelined exchange when the caller gives none: 4 for two ranks (measured), 2 beyond
        from msplat_b200.render import _default_grad_chunks, _resolve_group
        assert _default_grad_chunks(_resolve_group(True)) == 4
        assert _default_grad_chunks(lambda t: None) == 2
        q.put((rank, grads.flat.clone(), float(total)))
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference over all 6 views
    xyz, w = _make_params()
    grads = FlatGrads([xyz, w])
    cams = orbit_cameras(6)
    total = render_view_batch(lambda e: _loss(xyz, w, e), cams, 0, 1, grads)
    for rank, flat, _ in out:
        torch.testing.assert_close(flat, grads.flat, rtol=1e-5, atol=1e-6)
    assert abs(out[0][2] + out[1][2] - float(total)) < 1e-3 * abs(float(total))


def test_flat_grads_detached_views_are_caught():
    """zero_grad(set_to_none=True) detaches the .grad views from the flat buffer: all_reduce must not silently
    reduce stale zeros."""
    import pytest
    xyz, w = _make_params()
    grads = FlatGrads([xyz, w])
    (xyz.sum() + w.sum()).backward()
    assert grads.attached() and float(grads.flat.sum()) == xyz.numel() + w.numel()
    grads.all_reduce()  # world size 1: no collective, but the check runs
    xyz.grad = None     # what optimizer.zero_grad() does by default
    (xyz.sum() + w.sum()).backward()
    assert not grads.attached()
    with pytest.raises(RuntimeError, match="flat buffer"):
        grads.all_reduce()
    grads.attach()
    grads.zero_()
    (xyz.sum() + w.sum()).backward()
    grads.all_reduce()
    assert float(grads.flat.sum()) == xyz.numel() + w.numel()
