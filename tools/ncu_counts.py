#!/usr/bin/env python
"""Executed-operation counts of the fused env kernel from an `ncu --set full --import-source on` capture.

    python tools/ncu_counts.py gpurun_out/prof_rXX.ncu-rep ant_env_kernelILi0ELi1E 4096 > profiles/r2_counts_<Env>.json

Per SASS instruction the report holds how many THREADS executed it; classifying the opcode gives the executed
floating-point operations of the launch (FFMA = 2, FFMA2 = 4, FADD / FMUL = 1, FADD2 / FMUL2 = 2, MUFU = 1, DFMA = 2,
DADD / DMUL = 1) - an instrumented count of what the kernel really executes, redundant lanes and idle solver visits
included - and, with the per-line source mapping of tools/ncu_by_line.py, their split over the code regions.
bench.py reads the JSON for `roofline_fp32.executed_flop_per_env_step`, `roofline_issue` and `roofline.traffic`."""
import json
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from ncu_by_line import parse_disasm, parse_ncu  # noqa: E402
from ncu_report import regions  # noqa: E402

FP32 = {"FFMA": 2, "FFMA2": 4, "FADD": 1, "FMUL": 1, "FADD2": 2, "FMUL2": 2, "MUFU": 1, "FMNMX": 1, "FSEL": 0, "FSETP": 0}
FP64 = {"DFMA": 2, "DADD": 1, "DMUL": 1}


def opcode(sass):
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", sass)
    return m.group(1) if m else ""


def main():
    rep, tag, envs = sys.argv[1], sys.argv[2], int(sys.argv[3])
    so = os.path.join(ROOT, "hrl_pybullet_envs_b200", "libhrl_b200.so")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        open(os.path.join(tmp, "d.txt"), "w").write(subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout)
        open(os.path.join(tmp, "s.csv"), "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        D = parse_disasm(os.path.join(tmp, "d.txt"), tag)
        N = parse_ncu(os.path.join(tmp, "s.csv"))
    assert len(D) == len(N), "report and libhrl_b200.so are different builds (%d vs %d instructions)" % (len(D), len(N))
    import csv
    rows = list(csv.reader(raw.split("\n")))
    hdr, vals = rows[0], [r for r in rows if len(r) == len(rows[0])][-1]
    metric = {h: v for h, v in zip(hdr, vals)}
    R = regions()
    by_region = defaultdict(lambda: [0.0, 0.0, 0.0])   # fp32 flop, fp64 flop, warp instructions
    fp32 = fp64 = 0.0
    for i, (_, (f, ln), _) in enumerate(D):
        name = "other"
        for rf, lo, hi, rn in R:
            if f == rf and lo <= ln <= hi:
                name = rn
                break
        op = opcode(N[i]["sass"]).split(".")[0]
        a = FP32.get(op, 0) * N[i]["thr"]; b = FP64.get(op, 0) * N[i]["thr"]
        fp32 += a; fp64 += b
        by_region[name][0] += a; by_region[name][1] += b; by_region[name][2] += N[i]["inst"]
    winst = sum(n["inst"] for n in N)
    dram_unit = 1e6 if float(metric.get("dram__bytes_read.sum", 0) or 0) < 1e4 else 1.0   # the raw page prints Mbyte
    out = {"source": "ncu --set full of %s, kernel %s, %d envs per launch (tools/ncu_counts.py)" % (os.path.basename(rep), tag, envs),
           "envs_per_launch": envs,
           "warp_instructions_per_launch": winst,
           "executed_fp32_flop_per_launch": fp32, "executed_fp64_flop_per_launch": fp64,
           "executed_fp32_flop_per_env_step": fp32 / envs, "executed_fp64_flop_per_env_step": fp64 / envs,
           "dram_bytes_per_launch": (float(metric.get("dram__bytes_read.sum", 0) or 0) + float(metric.get("dram__bytes_write.sum", 0) or 0)) * dram_unit,
           "elapsed_cycles": float(metric.get("sm__cycles_elapsed.max", 0) or 0),
           "duration_us": float(metric.get("gpu__time_duration.sum", 0) or 0),
           "issue_active_pct": float(metric.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0) or 0),
           "regions": {k: {"fp32_flop_per_env_step": v[0] / envs, "fp64_flop_per_env_step": v[1] / envs, "warp_instructions": v[2]}
                       for k, v in sorted(by_region.items(), key=lambda kv: -kv[1][0])}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
