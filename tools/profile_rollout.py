"""A few fused rollouts and nothing else (ncu target: tools/ncu_report.py ... ant_env_kernelILi0ELi1ELi1E)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hrl_pybullet_envs_b200 import VecEnv

N, T = 4096, 16
env = VecEnv("AntGatherBulletEnv-v0", N, seed=0)
env.reset()
g = torch.Generator(device="cuda").manual_seed(7)
def lin(o, i, sc):
    return ((torch.rand(o, i, generator=g, device="cuda") * 2 - 1) * sc, (torch.rand(o, generator=g, device="cuda") * 2 - 1) * 0.1)
layers = (lin(64, env.D, 0.3), lin(64, 64, 0.2), lin(8, 64, 0.3))
buf = env.rollout_buffer(T)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):   # the first rollouts land the ants
    env.rollout_mlp(layers, T, buf=buf, sigma=0.1, noise_seed=i)
torch.cuda.synchronize()
print("ok", float(buf.rew.sum()))
