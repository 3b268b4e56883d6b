#!/usr/bin/env python
"""One-stop post-processing of an `ncu --set full --import-source on` capture of the env kernel.

    python tools/ncu_report.py gpurun_out/prof_rXX.ncu-rep [kernel-tag] > profiles/rXX_region_cycles.txt

Disassembles the in-tree libhrl_b200.so (nvdisasm -g carries the source lines), exports the SASS page of
the report, matches the two instruction by instruction (tools/ncu_by_line.py) and prints, per code region
of the kernel: static instructions, instructions actually executed, dynamic instructions per warp and the
share of the launch's cycles (stall samples x elapsed cycles / all samples).
"""
import os
import subprocess
import sys
import tempfile
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from ncu_by_line import parse_disasm, parse_ncu  # noqa: E402


def regions():
    """(file, first line, last line, name) from markers in the sources, so the table survives edits."""
    ant = open(os.path.join(ROOT, "hrl_pybullet_envs_b200/csrc/hrl_ant.cuh")).read().split("\n")
    cu = open(os.path.join(ROOT, "hrl_pybullet_envs_b200/csrc/hrl_b200.cu")).read().split("\n")

    def L(src, pat):
        return next(i for i, l in enumerate(src) if pat in l) + 1
    marks = [("fk", "LegKin leg_fk("), (None, "constexpr int sym6"), ("sinv_mul", "void sinv_mul("), (None, "float clampf("),
             ("emit_row", "void emit_row("), ("pgs_helpers", "Blackwell packed fp32"),
             ("ds_helpers", "// ---- Delassus-space sweep: helpers"), (None, "void ant_substep("),
             ("contacts_detect", "contacts: spheres vs"), ("dynamics", "smooth dynamics: bias"),
             ("rows_build", "constraint rows: counts"), ("pgs", "// ---------------- projected Gauss-Seidel, Bullet"),
             ("ds_sweep", "Delassus-space sweep (layout and derivation"), ("pgs_vel", "  float2 dv[7];"),
             ("integrate", "back to physical velocities")]
    pos = [(n, L(ant, p)) for n, p in marks] + [(None, len(ant) + 1)]
    out = [("hrl_ant.cuh", lo, hi - 1, n) for (n, lo), (_, hi) in zip(pos, pos[1:]) if n]
    wide = open(os.path.join(ROOT, "hrl_pybullet_envs_b200/csrc/hrl_ant_wide.cuh")).read().split("\n")
    wmarks = [("w_contact_helpers", "int add_cand_w("), ("w_dynamics", "void leg_dynamics_w("), ("w_emit_row", "void emit_row_w("),
              ("w_pgs_helpers", "struct RowW"), (None, "void ant_substep_w("), ("w_contacts_detect", "contacts: the leg's sphere slots"),
              ("w_dynamics_call", "smooth dynamics (replicated"), ("w_rows_build", "constraint rows: counts, visit positions; sub-lane"),
              ("w_pgs", "projected Gauss-Seidel, Bullet row order (see"), ("w_integrate", "back to physical velocities, clamp, integrate (as")]
    wpos = [(n, L(wide, p)) for n, p in wmarks] + [(None, len(wide) + 1)]
    out += [("hrl_ant_wide.cuh", lo, hi - 1, n) for (n, lo), (_, hi) in zip(wpos, wpos[1:]) if n]
    out += [("hrl_math.cuh", 1, 48, "math_helpers"), ("hrl_math.cuh", 49, 200, "rng"), ("hrl_sensors.cuh", 1, 400, "sensors")]
    a, b, c, d = L(cu, "// ---- load ----"), L(cu, "// ---- task layer: observation"), L(cu, "// ---- store ----"), L(cu, "// PointGather (point_bot.py")
    out += [("hrl_b200.cu", a - 20, b - 1, "load_physics_loop"), ("hrl_b200.cu", b, c - 1, "task_layer"), ("hrl_b200.cu", c, d, "store")]
    return out


def main():
    rep = sys.argv[1]
    tag = sys.argv[2] if len(sys.argv) > 2 else "ant_env_kernelILi0ELi1E"
    so = os.path.join(ROOT, "hrl_pybullet_envs_b200", "libhrl_b200.so")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
        open(os.path.join(tmp, "d.txt"), "w").write(dis)
        csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        open(os.path.join(tmp, "s.csv"), "w").write(csv)
        det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        D = parse_disasm(os.path.join(tmp, "d.txt"), tag)
        N = parse_ncu(os.path.join(tmp, "s.csv"))
    cyc = float(next(l for l in det.split("\n") if "Elapsed Cycles" in l).split()[-1].replace(",", ""))
    dur = next(l for l in det.split("\n") if "Duration" in l).split()[-2:]
    assert len(D) == len(N), "the report was taken with a different build of libhrl_b200.so (%d vs %d instructions)" % (len(D), len(N))
    R = regions()
    st = defaultdict(lambda: [0, 0, 0.0, 0.0, defaultdict(float)])
    tot = sum(n["samples"] for n in N)
    warps = max(n["inst"] for n in N if n["inst"] > 0 and n["inst"] == int(n["inst"]))  # prologue instructions: once per warp
    warps = N[0]["inst"] or warps
    for i, (_, (f, ln), _) in enumerate(D):
        name = "other"
        for rf, lo, hi, rn in R:
            if f == rf and lo <= ln <= hi:
                name = rn
                break
        s = st[name]
        s[0] += 1; s[1] += 1 if N[i]["inst"] > 0 else 0; s[2] += N[i]["inst"]; s[3] += N[i]["samples"]
        for kk, vv in N[i]["stalls"].items():
            s[4][kk[6:]] += vv
    print("%s: %s, kernel %s: %.0f elapsed cycles (%s %s), %d static instructions, %d warps" % (
        os.path.basename(rep), "ncu --set full", tag, cyc, dur[1], dur[0], len(D), warps))
    print("%-20s %7s %11s %10s %9s %6s  %s" % ("region", "static", "exec-static", "dyn/warp", "cyc/warp", "share", "top stall reasons (share of the region's samples)"))
    for k, v in sorted(st.items(), key=lambda kv: -kv[1][3]):
        top = sorted(v[4].items(), key=lambda kv: -kv[1])[:4]
        print("%-20s %7d %11d %10.0f %9.0f %5.1f%%  %s" % (k, v[0], v[1], v[2] / warps, v[3] * cyc / tot, 100 * v[3] / tot,
                                                         ", ".join("%s %.0f%%" % (a, 100 * b / max(v[3], 1)) for a, b in top)))
    allst = defaultdict(float)
    for v in st.values():
        for a, b in v[4].items():
            allst[a] += b
    print("all regions: " + ", ".join("%s %.1f%%" % (a, 100 * b / tot) for a, b in sorted(allst.items(), key=lambda kv: -kv[1])[:8]))
    print("total dyn/warp %.0f" % (sum(v[2] for v in st.values()) / warps))


if __name__ == "__main__":
    main()
