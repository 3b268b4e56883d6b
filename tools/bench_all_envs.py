"""Device-resident throughput of every env family at the BASELINE.json config sizes (exploration;
the headline metric is bench.py).  Same protocol as bench.py region 1: per-step CUDA-event pairs, L2 flushed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hrl_pybullet_envs_b200 import VecEnv

CASES = [("AntMjBulletEnv-v0", 4096, {}), ("AntGatherBulletEnv-v0", 4096, {}), ("AntMazeBulletEnv-v0", 4096, {}),
         ("AntFlagrunBulletEnv-v0", 16384, {}), ("AntMazeMjEnv-v0", 4096, {}), ("PointGatherBulletEnv-v0", 4096, {}),
         ("AntGatherBulletEnv-v0", 4096, dict(use_sensor=False)), ("AntMazeBulletEnv-v0", 4096, dict(sense_target=True)),
         ("AntFlagrunBulletEnv-v0", 4096, dict(use_sensor=True))]
K = int(sys.argv[1]) if len(sys.argv) > 1 else 400
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for env_id, N, kw in CASES:
    env = VecEnv(env_id, N, seed=0, **kw)
    env.reset()
    ring = torch.rand(16, N, env.A, device="cuda") * 2 - 1
    for i in range(200):
        env.step(ring[i % 16])
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize()
    for i in range(K):
        flush.zero_()
        ev[i][0].record(); env.step(ring[i % 16]); ev[i][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / K
    print("%-26s N=%-6d %-28s %7.1f us/step  %.3e env-steps/s" % (env_id, N, kw or "", ms * 1e3, N / (ms * 1e-3)))
    env.close()
