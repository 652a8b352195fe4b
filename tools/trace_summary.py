#!/usr/bin/env python3
"""Summarise a torch.profiler chrome trace (bench.py --trace): GPU kernels/memcpys in time order with the
idle gaps between them, per step -- where an end-to-end step loses time to the host."""
import json
import sys

d = json.load(open(sys.argv[1]))
ev = [e for e in d["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
busy_end = ev[0]["ts"]
gaps = []
for e in ev:
    if e["ts"] > busy_end + 20:  # > 20 us with nothing running on the device
        gaps.append((busy_end - t0, e["ts"] - busy_end, e["name"][:60]))
    busy_end = max(busy_end, e["ts"] + e["dur"])
total = busy_end - t0
print(f"{len(ev)} device activities over {total / 1e3:.2f} ms; idle gaps > 20 us:")
for at, dur, nxt in gaps:
    print(f"  at {at / 1e3:8.3f} ms  idle {dur:8.1f} us  before {nxt}")
print(f"idle total {sum(g[1] for g in gaps) / 1e3:.3f} ms")
if len(sys.argv) > 2:
    for e in ev:
        print(f"{(e['ts'] - t0) / 1e3:9.3f} {e['dur']:9.1f} us  s{e['args'].get('stream', '?')}  {e['name'][:80]}")
