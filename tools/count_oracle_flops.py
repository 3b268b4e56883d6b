#!/usr/bin/env python
"""Instrumented operation count of the f64 oracle -> coefficients of the algorithmic FLOP model.

The oracle is compiled as C++ with a `real` that counts every add / mul / div / sqrt / trig it executes
(oracle/flop_counter.hpp, `make -C oracle count`).  Batches of AntGather envs are stepped in regimes that separate
the terms of

    flops = A * env_substeps + B * rows + C * rows * solver_iterations + D * env_steps

(rows = 3 * contacts + joint-limit rows, summed over the env-substeps of the run): airborne vs landed ants, 0 / 2 / 5
solver iterations, 1 / 4 sub-steps per step.  A least-squares fit gives A..D; they are written to
profiles/r2_oracle_flop_model.json and hard-coded (with this provenance) in hrl_pybullet_envs_b200/roofline.py.
One FLOP = one add, mul, div, sqrt or trig call (a divide or square root is several machine operations; counting it
once keeps the figure a lower bound).  CPU only; runs in about a minute."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrl_pybullet_envs_b200.config import ENV_IDS  # noqa: E402
from oracle import oracle as O  # noqa: E402


def run(env_id, n, steps, settle, iters, substeps, seed=0):
    L = O.lib(count=True)
    cfg = O.default_config(ENV_IDS[env_id], n)
    cfg.seed = seed; cfg.solver_iters = iters; cfg.substeps = substeps
    env = O.OracleVecEnv(cfg, count=True)
    env.reset()
    rng = np.random.default_rng(seed)
    for _ in range(settle):
        env.step(rng.uniform(-1, 1, (n, env.A)).astype(np.float32))
    s0 = np.zeros(3); L.hrlo_stats(env.h, O._p(s0))
    L.hrlo_flop_counts_reset()
    for _ in range(steps):
        env.step(rng.uniform(-1, 1, (n, env.A)).astype(np.float32))
    c = (C.c_uint64 * 6)(); L.hrlo_flop_counts_get(c)
    s1 = np.zeros(3); L.hrlo_stats(env.h, O._p(s1))
    d = s1 - s0
    flops = float(sum(c))
    return dict(flops=flops, counts=[int(x) for x in c], contacts=d[0], limit_rows=d[1], env_substeps=d[2], env_steps=n * steps,
                rows=3 * d[0] + d[1], iters=iters)


def main():
    env_id = "AntGatherBulletEnv-v0"
    runs = []
    for iters in (0, 2, 5):
        for substeps in (1, 4):
            runs.append(run(env_id, 64, 6, 0, iters, substeps))        # airborne: joint-limit rows only
            runs.append(run(env_id, 64, 20, 60, iters, substeps))      # landed: contacts + limits
    X = np.array([[r["env_substeps"], r["rows"], r["rows"] * r["iters"], r["env_steps"]] for r in runs])
    y = np.array([r["flops"] for r in runs])
    coef, res, rank, _ = np.linalg.lstsq(X, y, rcond=None)
    pred = X @ coef
    err = np.abs(pred - y) / y
    A, B, Cc, D = [float(x) for x in coef]
    ref = run(env_id, 256, 50, 200, 5, 4, seed=1)                       # the bench's regime, as a check of the fit
    per_step = ref["flops"] / ref["env_steps"]
    Cbar = ref["contacts"] / ref["env_substeps"]; Lbar = ref["limit_rows"] / ref["env_substeps"]
    model = 4 * (A + (3 * Cbar + Lbar) * (B + 5 * Cc)) + D
    out = {"source": "tools/count_oracle_flops.py: counting C++ build of oracle/hrl_oracle.c (f64, 13-link ABA + ABA impulse response per row)",
           "per_env_substep": A, "per_row": B, "per_row_iteration": Cc, "per_env_step_task_layer": D,
           "fit_max_rel_error": float(err.max()),
           "check": {"regime": "256 AntGather envs, 200 settle + 50 counted steps, random actions",
                     "contacts_per_substep": Cbar, "limit_rows_per_substep": Lbar,
                     "counted_flop_per_env_step": per_step, "model_flop_per_env_step": model,
                     "op_mix": dict(zip(["add", "mul", "div", "sqrt", "trig", "fabs"], ref["counts"]))}}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_oracle_flop_model.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
