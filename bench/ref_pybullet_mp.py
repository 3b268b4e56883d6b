#!/usr/bin/env python
"""The REAL reference arm: unmodified `hrl_pybullet_envs` + pybullet in a multiprocessing vector env.

SURVEY.md 8(d) "CPU baseline": one worker process per host core, each worker owns one env created with
`gym.make(id)` exactly as the reference's README does (README.md:20-37) and steps it with U(-1,1) actions, resetting
when `done`.  Workers run free (no per-step lock-step barrier: the most favourable schedule for the CPU side); the
parent starts them together, lets every worker do `warmup` untimed steps, then times `steps` steps per worker.

    value = workers * steps / slowest worker's wall time        [env-steps/s on `workers` cores]

Runs iff pybullet, gym and the reference package are importable (`available()`); in this image they are not (no network,
no wheel), and `python bench/ref_pybullet_mp.py` prints "reference unavailable: pybullet not installed".  The
reference package is looked for as an installed module first, then under baseline/_ref (the offline pip target of the
bench contract), then under HRL_REFERENCE_PATH.
"""
import importlib.util
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_paths():
    return [p for p in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("HRL_REFERENCE_PATH")) if p and os.path.isdir(p)]


def why_unavailable():
    """'' when the real reference can run, else the first missing module."""
    for m in ("pybullet", "gym"):
        if importlib.util.find_spec(m) is None:
            return "module %r not found" % m
    if importlib.util.find_spec("hrl_pybullet_envs") is None:
        for p in _reference_paths():
            if os.path.isdir(os.path.join(p, "hrl_pybullet_envs")):
                return ""
        return "module 'hrl_pybullet_envs' not found"
    return ""


def available():
    return why_unavailable() == ""


def _worker(env_id, seed, warmup, steps, start_evt, out_q):
    for p in _reference_paths():
        if p not in sys.path:
            sys.path.insert(0, p)
    import gym
    import numpy as np
    import hrl_pybullet_envs  # noqa: F401  (registers the ids, hrl_pybullet_envs/__init__.py:9-16)
    env = gym.make(env_id)
    env.seed(seed)
    rng = np.random.RandomState(seed)
    env.reset()
    shape = env.action_space.shape
    out_q.put(("ready", seed))
    start_evt.wait()
    ep_ret, rets = 0.0, []
    t0 = None
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        _, rew, done, _ = env.step(rng.uniform(-1, 1, shape))   # README.md:30
        ep_ret += rew
        if done:
            rets.append(ep_ret); ep_ret = 0.0
            env.reset()
    el = time.perf_counter() - t0
    out_q.put(("done", seed, el, rets))


def run(env_id, steps, warmup, workers=None):
    workers = workers or len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")     # pybullet clients must not be forked
    q, go = ctx.Queue(), ctx.Event()
    ps = [ctx.Process(target=_worker, args=(env_id, s, warmup, steps, go, q), daemon=True) for s in range(workers)]
    for p in ps:
        p.start()
    for _ in ps:
        assert q.get(timeout=600)[0] == "ready"
    go.set()
    res = [q.get(timeout=3600) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    slowest = max(r[2] for r in res)
    rets = [x for r in res for x in r[3]]
    return {"value": workers * steps / slowest, "ms_per_step": 1e3 * slowest / max(steps, 1), "workers": workers,
            "episodes": len(rets), "mean_return": (sum(rets) / len(rets)) if rets else None,
            "sample": "%d pybullet envs (one per worker process) x %d steps after %d warm-up steps" % (workers, steps, warmup)}


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="AntGatherBulletEnv-v0")
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--workers", type=int, default=None)
    a = ap.parse_args()
    why = why_unavailable()
    if why:
        print("reference unavailable: pybullet not installed (%s)" % why)
        return 0
    print(json.dumps(run(a.env, a.steps, a.warmup, a.workers)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
