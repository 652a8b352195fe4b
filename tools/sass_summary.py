#!/usr/bin/env python3
"""Opcode histogram per kernel of the in-tree library (`cuobjdump -sass msplat_b200/libmsplat_b200.so`): the evidence
for the sm_100a features the kernels claim (TMA bulk copies UBLKCP / UBLKRED, mbarrier SYNCS, cp.async LDGSTS, packed
FP32 FFMA2 / FMUL2 / FADD2, MUFU, vector reductions RED, warp votes).  Writes profiles/sass_summary.txt.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "msplat_b200", "libmsplat_b200.so")
KEY = ["UBLKCP", "UBLKRED", "SYNCS", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "MUFU", "RED", "ATOMG", "ATOMS", "VOTE", "SHFL",
       "LDG", "STG", "LDS", "STS", "FFMA", "FMUL", "FADD", "DFMA", "DADD", "DMUL", "F2I", "BAR", "UTMALDG", "UTCMMA", "HMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = kern.replace("(int)", "").replace("(bool)", "")
            kern = re.sub(r"\(.*", "", kern).replace("void ", "")
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            hist[kern][m.group(1)] += 1
    total = collections.Counter()
    for c in hist.values():
        total.update(c)
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (sm_100a), {len(hist)} kernels, "
          f"{sum(total.values())} instructions")
    print("# library totals: " + ", ".join(f"{k} {total[k]}" for k in KEY if total[k]))
    print("# (no UTMALDG / UTCMMA / HMMA: the slabs are 1-D bulk copies and no stage is a dense contraction)\n")
    # merge template instantiations of the same kernel family for the table; list the hot instantiations in full
    fam = collections.OrderedDict()
    for k, c in hist.items():
        base = re.sub(r"<.*", "", k)
        fam.setdefault(base, [0, collections.Counter()])
        fam[base][0] += 1
        fam[base][1].update(c)
    print(f"{'kernel family':42s} inst  " + " ".join(f"{k:>7s}" for k in KEY[:17]))
    for base, (n, c) in fam.items():
        print(f"{base[:42]:42s} {n:4d}  " + " ".join(f"{c[k]:7d}" for k in KEY[:17]))
    print("\n# hot instantiations (BASELINE config #3: SH degree 3, C = 4)")
    hot = ["render_pre_fwd_kernel<3>", "render_pre_bwd_kernel<3, false>", "blend_fwd_kernel<4, false>",
           "blend_bwd_kernel<4, 256, 3>", "onesweep_kernel<false, 8, 8, 16>", "onesweep_kernel<true, 8, 8, 16>",
           "keygen_kernel", "duplicate_kernel", "scan_offsets_kernel", "tile_range_kernel", "adam_kernel"]
    for k, c in hist.items():
        if any(k.replace("msb::", "") == h for h in hot):
            top = ", ".join(f"{op} {n}" for op, n in c.most_common(14))
            print(f"{k}: {sum(c.values())} instructions: {top}")


if __name__ == "__main__":
    main()
