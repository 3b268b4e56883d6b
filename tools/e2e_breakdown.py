"""Where the end-to-end (host-buffer) step time goes: kernel+sync, raw C-ABI zero-copy / copy, VecEnv.step(numpy)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hrl_pybullet_envs_b200 import VecEnv, _cabi

N, K = 4096, 500
env = VecEnv("AntGatherBulletEnv-v0", N, seed=0)
env.reset()
ring = torch.rand(64, N, 8, device="cuda") * 2 - 1
ring_h = ring.cpu().pin_memory().numpy()
for i in range(200):
    env.step(ring[i % 64])
torch.cuda.synchronize()

def timeit(f, k=K):
    for i in range(10): f(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(k): f(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e6

def dev_sync(i):
    env.step(ring[i % 64]); torch.cuda.synchronize()
print("device step + sync per step        %.1f us" % timeit(dev_sync))
H = env._host_buffers(); S = H["sets"][0]; st = env._stream()
def raw(i):
    _cabi.check(env.L.hrl_step_host(env.h, H["p_act"], S["p_obs"], S["p_rew"], S["p_done"], S["p_info"], st))
for mode in ("zerocopy", "copy"):
    env.set_host_mode(mode)
    print("raw hrl_step_host %-8s          %.1f us" % (mode, timeit(raw)))
# zero-copy with outputs only partially on the host: obs stays on the device
env.set_host_mode("auto")
def vec(i):
    env.step(ring_h[i % 64])
print("VecEnv.step(numpy) auto            %.1f us" % timeit(vec))
def cp(i):
    H["act_np"][...] = ring_h[i % 64]
print("numpy action copy only             %.1f us" % timeit(cp))
