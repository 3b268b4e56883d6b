"""Step one env family for an ncu capture:  ncu ... python tools/profile_env.py ENV_ID [N] [STEPS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hrl_pybullet_envs_b200 import VecEnv
env_id = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096; T = int(sys.argv[3]) if len(sys.argv) > 3 else 230
env = VecEnv(env_id, N, seed=0); env.reset()
ring = torch.rand(16, N, env.A, device="cuda") * 2 - 1
for t in range(T):
    env.step(ring[t % 16])
torch.cuda.synchronize()
