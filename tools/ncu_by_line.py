#!/usr/bin/env python
"""Attribute an ncu SASS-level source page to CUDA source lines / code regions.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:ant_env --launch-count 1 > sass.csv
    cuobjdump -xelf all libhrl_b200.so && nvdisasm -g -c hrl_b200.sm_100a.cubin > disasm.txt
    python tools/ncu_by_line.py sass.csv disasm.txt ant_env_kernelILi0

The i-th instruction of the kernel in the ncu page is matched with the i-th instruction of the
nvdisasm listing (which carries //## File ... line ... annotations, innermost inlined frame).
"""
import csv
import re
import sys
from collections import defaultdict


def parse_disasm(path, kernel_tag):
    lines = open(path).read().split("\n")
    start = None
    for i, l in enumerate(lines):
        if l.startswith(".text.") and kernel_tag in l and l.endswith(":"):
            start = i
            break
    out = []
    cur = ("?", 0)
    for l in lines[start + 1:]:
        if l.startswith("//---") or l.startswith("\t.section"):
            if out:
                break
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def parse_ncu(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[hi + 1:]:
        if r and r[0] in ("Kernel Name", "Address"):
            break  # next launch
        if len(r) < len(hdr):
            continue
        out.append({"sass": r[ci["Source"]], "samples": float(r[ci["# Samples"]] or 0),
                    "inst": float(r[ci["Instructions Executed"]] or 0),
                    "thr": float(r[ci["Thread Instructions Executed"]] or 0),
                    "stalls": {h: float(r[i] or 0) for h, i in ci.items() if h.startswith("stall_") and "Not Issued" not in h}})
    return out


def main():
    sass_csv, disasm, tag = sys.argv[1:4]
    regions = []
    if len(sys.argv) > 4:  # file:lo-hi=name ...
        for spec in sys.argv[4:]:
            loc, name = spec.split("=")
            f, rng = loc.split(":")
            lo, hi = map(int, rng.split("-"))
            regions.append((f, lo, hi, name))
    D = parse_disasm(disasm, tag)
    N = parse_ncu(sass_csv)
    print("instructions: disasm %d, ncu %d" % (len(D), len(N)))
    n = min(len(D), len(N))
    by_line = defaultdict(lambda: [0.0, 0.0, 0.0])
    by_region = defaultdict(lambda: [0.0, 0.0, 0.0, defaultdict(float)])
    tot_i = tot_s = 0.0
    for i in range(n):
        (f, ln) = D[i][1]
        a = by_line[(f, ln)]
        a[0] += N[i]["inst"]; a[1] += N[i]["samples"]; a[2] += N[i]["thr"]
        tot_i += N[i]["inst"]; tot_s += N[i]["samples"]
        name = "other"
        for (rf, lo, hi, rn) in regions:
            if f == rf and lo <= ln <= hi:
                name = rn
                break
        b = by_region[name]
        b[0] += N[i]["inst"]; b[1] += N[i]["samples"]; b[2] += N[i]["thr"]
        for k, v in N[i]["stalls"].items():
            b[3][k] += v
    print("total warp-instructions %.0f, samples %.0f" % (tot_i, tot_s))
    if regions:
        print("\n%-28s %12s %7s %9s %7s %6s  top stalls" % ("region", "warp-inst", "share", "samples", "share", "thr/in"))
        for name, (ins, smp, thr, st) in sorted(by_region.items(), key=lambda kv: -kv[1][1]):
            top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
            print("%-28s %12.0f %6.1f%% %9.0f %6.1f%% %6.1f  %s" % (name, ins, 100 * ins / tot_i, smp, 100 * smp / tot_s, thr / max(ins, 1),
                                                                   ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(smp, 1)) for k, v in top)))
    print("\ntop source lines by stall samples")
    for (f, ln), (ins, smp, thr) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-18s %5d  inst %10.0f (%4.1f%%)  samples %8.0f (%4.1f%%)" % (f, ln, ins, 100 * ins / tot_i, smp, 100 * smp / tot_s))


if __name__ == "__main__":
    main()
