#!/usr/bin/env python3
"""BASELINE config #2 acceptance: the reference's 2-D fitting tutorial at its own shape --
data/stanford-bunny.jpg 512x512 (tests/golden/gs2d_target.png), 100k Gaussians, seed 123, Adam lr 0.01, SmoothL1,
500 steps -- with the unmodified reference build (torch.optim.Adam) and with msplat_b200 (fused Adam and
torch.optim.Adam): loss at initialisation and after 500 steps, device time per iteration, and how far the loss
curves are apart compared with two runs of the reference itself (its backward is not deterministic).

    python tools/gs2d_acceptance.py [--points 100000] [--iters 500] [--out gpurun_out/gs2d_acceptance.json]
"""
import argparse
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

import torch  # noqa: E402


def tutorial():
    spec = importlib.util.spec_from_file_location("gs_2d", os.path.join(ROOT, "tutorials", "gs_2d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=100000)
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gs2d_acceptance.json"))
    args = ap.parse_args()
    import msplat_b200
    try:
        import msplat as ref
    except Exception:
        ref = None
    t = tutorial()
    target = t.load_target(None, 512).cuda()
    runs = {}

    def run(name, api, optimizer):
        timing = {}
        losses = t.fit(api, target, args.points, args.iters, quiet=True, optimizer=optimizer, timing=timing)
        runs[name] = {"losses": losses, "ms_per_iteration": timing["ms_per_iteration"]}
        print(f"{name:28s} loss {losses[0]:.6f} -> {losses[-1]:.6f}  {timing['ms_per_iteration']:.3f} ms/iteration", flush=True)

    run("msplat_b200 fused Adam", msplat_b200, "fused")
    run("msplat_b200 torch Adam", msplat_b200, "torch")
    if ref is not None:
        run("reference run 1", ref, "torch")
        run("reference run 2", ref, "torch")
    marks = [k for k in (0, 10, 50, 100, 200, 300, 400, args.iters - 1) if k < args.iters]
    out = {"points": args.points, "iters": args.iters, "target": "tests/golden/gs2d_target.png (512x512)",
           "runs": {n: {"loss_at": {str(k): r["losses"][k] for k in marks}, "ms_per_iteration": r["ms_per_iteration"]}
                    for n, r in runs.items()}}
    if ref is not None:
        a, b = runs["reference run 1"]["losses"], runs["reference run 2"]["losses"]
        o = runs["msplat_b200 fused Adam"]["losses"]
        rel = lambda x, y: [abs(p - q) / abs(q) for p, q in zip(x, y)]
        rr, orr = rel(b, a), rel(o, a)
        out["relative_loss_difference"] = {
            "reference_vs_reference_max": {str(k): max(rr[:k + 1]) for k in marks},
            "ours_vs_reference_max": {str(k): max(orr[:k + 1]) for k in marks}}
        out["speedup_per_iteration"] = runs["reference run 1"]["ms_per_iteration"] / runs["msplat_b200 fused Adam"]["ms_per_iteration"]
        print(json.dumps(out["relative_loss_difference"], indent=1))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    # full curves for the record (small)
    json.dump({n: r["losses"] for n, r in runs.items()}, open(args.out.replace(".json", "_curves.json"), "w"))


if __name__ == "__main__":
    main()
