#!/usr/bin/env python3
"""Summarise ncu outputs for profiles/: `ncu_summary.py launches <csv> <marker-regex>` prints the
per-kernel share table of the launch list (the launches between the last-but-N marker kernels);
`ncu_summary.py full <raw.csv>` prints one row per captured launch of an `ncu --page raw --csv` dump."""
import csv
import re
import sys


def launches(path, first="render_pre_fwd", last="render_pre_bwd", which=3):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    L = []
    for r in rows[1:]:
        try:
            L.append((r[ki], float(r[vi].replace(",", ""))))
        except ValueError:
            pass
    idx = [i for i, (k, _) in enumerate(L) if first in k]
    s = idx[which] if len(idx) > which else idx[-1]
    e = [i for i, (k, _) in enumerate(L) if last in k and i > s][0]
    step = L[s:e + 1]
    tot = sum(v for _, v in step)
    agg = {}
    for k, v in step:
        k = re.sub(r"\(.*", "", k).replace("void ", "").replace("msb::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    print(f"One fwd+bwd render (step {which + 1} of the command = first timed step): {len(step)} launches, "
          f"{tot / 1e3:.1f} us of kernel time\n")
    print("| kernel | launches | us | share |\n|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {n} | {v / 1e3:.1f} | {100 * v / tot:.1f}% |")
    ours = sum(v for k, (n, v) in agg.items() if not k.startswith("at::"))
    print(f"\nOur kernels: {100 * ours / tot:.1f}% of kernel time.")


def full(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    scale = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0,
             "Gbyte": 1e3}
    cols = [("us", "gpu__time_duration.sum"), ("DRAM rd MB", "dram__bytes_read.sum"), ("DRAM wr MB", "dram__bytes_write.sum"),
            ("DRAM %", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
            ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
            ("regs", "launch__registers_per_thread"), ("warp insts", "smsp__inst_executed.sum"),
            ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio")]
    cols = [(a, b) for a, b in cols if b in hdr]
    print("| kernel | " + " | ".join(a for a, _ in cols) + " |\n|---|" + "---|" * len(cols))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "")
        vals = []
        for _, b in cols:
            v = r[hdr.index(b)].replace(",", "")
            try:
                vals.append(f"{float(v) * scale.get(units[hdr.index(b)], 1.0):.4g}")
            except ValueError:
                vals.append(v)
        print(f"| `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], which=int(sys.argv[3]) if len(sys.argv) > 3 else 3)
    else:
        full(sys.argv[2])
