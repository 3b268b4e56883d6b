"""debug helper: which observation component differs most between CUDA and the oracle in the one-step test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hrl_pybullet_envs_b200 import VecEnv
from oracle import oracle as O
env_id = sys.argv[1] if len(sys.argv) > 1 else "AntGatherBulletEnv-v0"
N, T = 512, 120
g = VecEnv(env_id, N, seed=11); o = O.OracleVecEnv.make(env_id, N, seed=11)
g.reset(); o.reset()
gen = torch.Generator().manual_seed(1)
for t in range(T):
    a = (torch.rand(N, g.A, generator=gen) * 2 - 1)
    if t % 4 == 0:
        f, i = g.get_state()
        o.set_state(f.cpu().numpy().astype(np.float64), i.cpu().numpy())
        og, rg, dg, info = g.step(a.cuda())
        oo, ro, do, io = o.step(a.numpy())[:4]
        og = og.cpu().numpy(); dg = dg.cpu().numpy()
        live = (dg == do) & ~dg
        err = np.abs(og - oo) * live[:, None]
        e, c = np.unravel_index(err.argmax(), err.shape)
        if err.max() > 5e-3:
            f2, _ = g.get_state(); fo, _ = o.get_state()
            print("t", t, "env", e, "col", c, "gpu", og[e, c], "ora", oo[e, c], "state err", np.abs(f2.cpu().numpy()[e] - fo[e]).max())
            print("  quat", f[e, 3:7].cpu().numpy() if f.shape[1] > 7 else None)
    else:
        g.step(a.cuda())
print("done")
