"""Long random-action soak of every env family on the GPU: finite observations / rewards, states inside the arena,
joint angles near their limits, episode bookkeeping consistent.  python tools/soak.py [steps] [envs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hrl_pybullet_envs_b200 import VecEnv, config as K

T = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
LO = torch.tensor([-0.698132, 0.523599, -0.698132, -1.745329, -0.698132, -1.745329, -0.698132, 0.523599], device="cuda")
HI = torch.tensor([0.698132, 1.745329, 0.698132, -0.523599, 0.698132, -0.523599, 0.698132, 1.745329], device="cuda")
CASES = [("AntGatherBulletEnv-v0", {}, 7.5), ("AntGatherBulletEnv-v0", dict(robot_coll_dist=0), 7.5), ("AntMazeBulletEnv-v0", {}, 9.0),
         ("AntMazeMjEnv-v0", {}, 9.0), ("AntFlagrunBulletEnv-v0", {}, 6.0), ("AntMjBulletEnv-v0", {}, 1e9),
         ("PointGatherBulletEnv-v0", {}, 7.5), ("PointGatherBulletEnv-v0", dict(use_sensor=False), 7.5)]
for env_id, kw, half in CASES:
    env = VecEnv(env_id, N, seed=123, **kw)
    obs = env.reset()
    g = torch.Generator(device="cuda").manual_seed(7)
    ndone = 0; ret = torch.zeros(N, device="cuda"); worst_q = 0.0; worst_xy = 0.0
    for t in range(T):
        # mostly random actions; every so often a stretch of saturated ones (drives joints into their limits, ants into walls)
        a = torch.rand(N, env.A, generator=g, device="cuda") * 2 - 1
        if (t // 50) % 5 == 4:
            a = torch.sign(a)
        obs, rew, done, info = env.step(a)
        ndone += int(done.sum())
        if t % 100 == 99 or t == T - 1:
            assert torch.isfinite(obs).all(), (env_id, t, "obs")
            assert torch.isfinite(rew).all(), (env_id, t, "rew")
            f, i = env.get_state()
            assert torch.isfinite(f[:, :K.SF_ITEMS]).all(), (env_id, t, "state")
            worst_xy = max(worst_xy, float(f[:, K.SF_POS:K.SF_POS + 2].abs().max()))
            if env.kind != K.HRL_POINT_GATHER:
                q = f[:, K.SF_Q:K.SF_Q + 8]
                worst_q = max(worst_q, float(torch.maximum(LO - q, q - HI).max()))
                quat = f[:, K.SF_QUAT:K.SF_QUAT + 4]
                assert (quat.norm(dim=1) - 1).abs().max() < 1e-4
            assert (i[:, K.SI_T] <= 2000).all() and (i[:, K.SI_T] >= 0).all()
    assert worst_xy < half + 0.5, (env_id, worst_xy)
    assert worst_q < 1.0, (env_id, worst_q)      # soft limits (ERP 0.2, 5 iterations, limit rows visited before the contacts): saturated torques push up to ~0.6 rad past
    st = env.episode_stats()
    print("%-26s %-28s %d steps x %d envs ok: episodes %d, |xy| max %.2f, joint overshoot max %.3f rad, mean return %.3f" % (
        env_id, kw or "", T, N, ndone, worst_xy, worst_q, st["mean_return"]))
    env.close()
