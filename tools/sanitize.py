"""Tiny run of every env family for compute-sanitizer (memcheck / initcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hrl_pybullet_envs_b200 import VecEnv
for env_id, kw in [("AntGatherBulletEnv-v0", {}), ("AntGatherBulletEnv-v0", dict(use_sensor=False)), ("AntMazeBulletEnv-v0", dict(sense_target=True)),
                   ("AntFlagrunBulletEnv-v0", dict(use_sensor=True)), ("AntMjBulletEnv-v0", {}), ("AntMazeMjEnv-v0", {}), ("PointGatherBulletEnv-v0", {})]:
    env = VecEnv(env_id, 43, seed=1, max_episode_steps=12, **kw)   # 43: a ragged last warp; short episodes: resets inside the run
    env.reset()
    for t in range(30):
        env.step(torch.rand(43, env.A, device="cuda") * 2 - 1, want_terminal_obs=True)
    import numpy as np
    env.step(np.random.uniform(-1, 1, (43, env.A)).astype(np.float32))
    torch.cuda.synchronize()
    print("ok", env_id, kw)
