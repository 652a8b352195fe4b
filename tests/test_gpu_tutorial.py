"""tutorials/gs_2d.py (BASELINE config #2's application): the optimisation loop of
the reference's 2-D fitting tutorial runs through the drop-in API, the loss falls, and the loss
curve tracks the unmodified reference build's on the same seed."""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tutorial():
    spec = importlib.util.spec_from_file_location("gs_2d", os.path.join(ROOT, "tutorials", "gs_2d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_gs2d_loss_falls():
    import msplat_b200
    t = _tutorial()
    target = t.procedural_target(128, 128).cuda()
    losses = t.fit(msplat_b200, target, points=4000, iters=150, quiet=True)
    assert losses[-1] < 0.5 * losses[0], f"loss did not fall: {losses[0]} -> {losses[-1]}"
    assert all(l == l for l in losses), "NaN in the loss curve"


def test_gs2d_tracks_reference(ref_msplat):
    import msplat_b200
    t = _tutorial()
    target = t.procedural_target(96, 96).cuda()
    a = t.fit(msplat_b200, target, points=3000, iters=12, quiet=True)
    b = t.fit(ref_msplat, target, points=3000, iters=12, quiet=True)
    for k, (x, y) in enumerate(zip(a, b)):
        assert abs(x - y) <= 2e-3 * abs(y) + 1e-6, f"iteration {k}: {x} vs reference {y}"
