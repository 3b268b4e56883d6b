"""Per-warp duration of the fused Ant step (profiling build -DHRL_WARP_TIMES): which warps are the slow ones, and why.
The launch time at 4096 envs is the time of the SLOWEST of the 512 warps (one per scheduler), so the tail of this
distribution - not its mean - is what the step costs.  Usage (GPU box):
    python tools/warp_times.py [env-id] [num-envs]
builds hrl_pybullet_envs_b200/libhrl_b200_wt.so beside the product library and runs against it."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib_wt = os.path.join(ROOT, "hrl_pybullet_envs_b200", "libhrl_b200_wt.so")
os.environ["HRL_B200_LIB"] = lib_wt
from hrl_pybullet_envs_b200 import _cabi  # noqa: E402

_cabi.build(defines={"HRL_WARP_TIMES": 1}, out=lib_wt)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from hrl_pybullet_envs_b200 import VecEnv  # noqa: E402

env_id = sys.argv[1] if len(sys.argv) > 1 else "AntGatherBulletEnv-v0"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
env = VecEnv(env_id, N, seed=0)
env.reset()
g = torch.Generator(device="cuda").manual_seed(1000)
ring = torch.rand(64, N, env.A, generator=g, device="cuda") * 2 - 1
for i in range(300):
    env.step(ring[i % 64])
W = (N + 7) // 8
L = env.L
L.hrl_debug_warp_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
recs = []
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for i in range(200):
    flush.zero_()
    _, _, done, _ = env.step(ring[i % 64])
    out = np.zeros((W, 4), dtype=np.uint64)
    _cabi.check(L.hrl_debug_warp_times(env.h, out.ctypes.data_as(C.c_void_p), W))
    recs.append(out.astype(np.int64))
R = np.stack(recs)                      # [steps, warps, 4]
cyc, phys, trips, misc = R[..., 0], R[..., 1], R[..., 2], R[..., 3]
nl, nc = trips & 0xffff, trips >> 16
passes, resets = misc & 0xff, misc >> 8
mx = cyc.max(axis=1)
print("%s, %d envs, %d warps, 200 steps (L2 flushed)" % (env_id, N, W))
print("per-step cycles of a warp: mean %.0f  median %.0f  p90 %.0f  p99 %.0f | slowest warp of a step: mean %.0f  (min %d, max %d)" % (
    cyc.mean(), np.median(cyc), np.percentile(cyc, 90), np.percentile(cyc, 99), mx.mean(), mx.min(), mx.max()))
print("physics loop share of a warp's cycles: %.1f%%; task layer + load + store: mean %.0f cycles" % (100 * phys.mean() / cyc.mean(), (cyc - phys).mean()))
print("solver trips per step (sum over 4 sub-steps of the max over the warp's 8 envs): limits mean %.1f, contacts mean %.1f" % (nl.mean(), nc.mean()))
has_reset = resets > 0
print("warps with a reset in the step: %.1f%%; cycles with / without a reset: %.0f / %.0f; task passes mean %.2f" % (
    100 * has_reset.mean(), cyc[has_reset].mean() if has_reset.any() else 0, cyc[~has_reset].mean(), passes.mean()))
# least-squares model of a warp's cycles
X = np.stack([np.ones(cyc.size), nl.ravel(), nc.ravel(), has_reset.ravel().astype(float)], 1)
coef, *_ = np.linalg.lstsq(X, cyc.ravel().astype(float), rcond=None)
print("cycles ~ %.0f + %.0f * limit-trips + %.0f * contact-trips + %.0f * [reset in warp]   (residual std %.0f)" % (
    coef[0], coef[1], coef[2], coef[3], (cyc.ravel() - X @ coef).std()))
# who is the slowest warp?
am = cyc.argmax(axis=1)
sl_nc = nc[np.arange(len(am)), am]; sl_nl = nl[np.arange(len(am)), am]; sl_rs = has_reset[np.arange(len(am)), am]
print("slowest warp of each step: contact-trips mean %.1f (all warps %.1f), limit-trips %.1f (%.1f), has a reset in %.0f%% of the steps" % (
    sl_nc.mean(), nc.mean(), sl_nl.mean(), nl.mean(), 100 * sl_rs.mean()))
for q in (50, 90, 99, 100):
    print("  contact-trips p%d = %d, limit-trips p%d = %d" % (q, np.percentile(nc, q), q, np.percentile(nl, q)))
