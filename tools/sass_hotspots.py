#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample shares of one kernel of an `ncu --set full --import-source on`
report: joins ncu's SASS page with `nvdisasm -g` of the in-tree library (the same build that was profiled).

    python tools/sass_hotspots.py gpurun_out/prof.ncu-rep blend_bwd [top=40]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                          "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    name = [r[1] for r in rows if r and r[0] == "Kernel Name"][0]
    his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr = rows[his[0]]
    end = his[1] - 1 if len(his) > 1 else len(rows)
    ii, st, si = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
    prof = []
    for r in rows[his[0] + 1:end]:
        if len(r) == len(hdr):
            try:
                prof.append((float(r[ii]), float(r[st]), r[si].strip()))
            except ValueError:
                pass
    # the cubin function whose demangled name matches
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "msplat_b200", "libmsplat_b200.so")], cwd=tmp,
                   capture_output=True)
    norm = lambda n: re.sub(r"\(.*", "", n.replace("(int)", "").replace("(bool)", "")).replace("void ", "").replace(" ", "")
    short = norm(name)
    best = None
    for f in os.listdir(tmp):
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if not txt:
            continue
        lines = txt.split("\n")
        for i, l in enumerate(lines):
            m = re.match(r"^\.text\.(\S+):", l)
            if not m:
                continue
            dem = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            if norm(dem) == short:
                ins, cur = [], ("?", 0)
                for l2 in lines[i + 1:]:
                    if l2.startswith(".text.") and "L_x" not in l2 and ins:
                        break
                    m2 = re.match(r'\s*//## File "([^"]+)", line (\d+)', l2)
                    if m2:
                        cur = (m2.group(1).split("/")[-1], int(m2.group(2)))
                        continue
                    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l2):
                        ins.append(cur)
                    if l2.startswith("//-----") and ins:
                        break
                best = ins
                break
        if best:
            break
    assert best, "kernel not found in the library"
    n = min(len(best), len(prof))
    tot, tots = sum(p[0] for p in prof), max(sum(p[1] for p in prof), 1.0)
    by, bs = collections.Counter(), collections.Counter()
    for k in range(n):
        by[best[k]] += prof[k][0]
        bs[best[k]] += prof[k][1]
    print(f"{short}: {len(prof)} SASS instructions profiled, {len(best)} disassembled; {tot:.3e} warp instructions")
    src = {}
    for (f, l), v in by.most_common(top):
        try:
            if f not in src:
                src[f] = open(os.path.join(ROOT, "msplat_b200", "csrc", f)).read().split("\n")
            line = src[f][l - 1].strip()[:110]
        except Exception:
            line = ""
        print(f"{100 * v / tot:5.1f}% insts {100 * bs[(f, l)] / tots:5.1f}% stalls  {f}:{l}  {line}")


if __name__ == "__main__":
    main()
