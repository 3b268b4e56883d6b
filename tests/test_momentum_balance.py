"""Momentum balance in free flight - a property of the physics that needs no reference data.

Ants thrown into the air (no contact, torques off, link damping off) are a closed system under gravity: over a time T
the total linear momentum changes by  M_total * g * T  and the angular momentum about the common centre of mass stays
put (joint-limit rows are internal forces and may fire).  The semi-implicit Euler step of Bullet (velocities from
M(q_old), then positions) satisfies both only to first order in the sub-step h, so the test checks the size of the
residual over one env step AND its convergence: 16 sub-steps of h / 4 must leave less than half the residual of 4
sub-steps of h (first order: a quarter).  Both momenta are evaluated with the INDEPENDENT numpy model of
tests/test_oracle_lagrangian.py (explicit forward kinematics from assets/ant.xml, per-link Jacobians), not with anything
the oracle or the kernels compute:
  P = sum_i m_i Jv_i u,   L = sum_i  I_i Jw_i u + m_i (c_i - c) x Jv_i u.
The oracle runs on the CPU; the CUDA path runs the BASELINE batch (4096 envs) and is checked on a sample of it."""
import numpy as np
import pytest

from hrl_pybullet_envs_b200 import config as K
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_oracle_lagrangian import links, _quat_R, M_TORSO, M_S, M_L

M_TOTAL = M_TORSO + 4 * (2 * M_S + M_L)
DT, G = 0.0165, 9.8


def momenta(f):
    """(P[3], L_com[3]) of one env from a row of the public state layout (HRL_SF_*)."""
    pos, R, q = f[K.SF_POS:K.SF_POS + 3].astype(float), _quat_R(f[K.SF_QUAT:K.SF_QUAT + 4].astype(float)), f[K.SF_Q:K.SF_Q + 8].astype(float)
    u = np.concatenate([f[K.SF_ANGVEL:K.SF_ANGVEL + 3], f[K.SF_LINVEL:K.SF_LINVEL + 3], f[K.SF_QD:K.SF_QD + 8]]).astype(float)
    ls = links(pos, R, q)
    com = sum(m * c for m, _, c, _, _ in ls) / M_TOTAL
    P = sum(m * (Jv @ u) for m, _, _, _, Jv in ls)
    L = sum(I @ (Jw @ u) + m * np.cross(c - com, Jv @ u) for m, I, c, Jw, Jv in ls)
    return P, L


def thrown_states(f, rng):
    n = f.shape[0]
    qt = rng.normal(size=(n, 4)); qt /= np.linalg.norm(qt, axis=1, keepdims=True)
    f[:, K.SF_POS:K.SF_POS + 3] = np.c_[rng.uniform(-1, 1, (n, 2)), np.full(n, 3.0)]
    f[:, K.SF_QUAT:K.SF_QUAT + 4] = qt
    lo = np.array([-0.5, 0.7, -0.5, -1.5, -0.5, -1.5, -0.5, 0.7]); hi = np.array([0.5, 1.5, 0.5, -0.7, 0.5, -0.7, 0.5, 1.5])
    f[:, K.SF_Q:K.SF_Q + 8] = rng.uniform(lo, hi, (n, 8))
    f[:, K.SF_LINVEL:K.SF_LINVEL + 3] = rng.normal(size=(n, 3))
    f[:, K.SF_ANGVEL:K.SF_ANGVEL + 3] = 2 * rng.normal(size=(n, 3))
    f[:, K.SF_QD:K.SF_QD + 8] = 2 * rng.normal(size=(n, 8))
    return f


def residuals(f0, f1, idx):
    """mean over the sampled envs of |dP - M g T| / (M g T) and of |dL| / max(|L|, 1)."""
    rp, rl = [], []
    for j in idx:
        P0, L0 = momenta(f0[j]); P1, L1 = momenta(f1[j])
        rp.append(np.abs(P1 - P0 - np.array([0, 0, -M_TOTAL * G * DT])).max() / (M_TOTAL * G * DT))
        rl.append(np.abs(L1 - L0).max() / max(np.abs(L0).max(), 1.0))
    return float(np.mean(rp)), float(np.mean(rl))


def judge(coarse, fine, who):
    (p4, l4), (p16, l16) = coarse, fine
    print("%s: linear momentum residual %.2e -> %.2e of the gravity impulse, angular momentum drift %.2e -> %.2e (h -> h / 4)"
          % (who, p4, p16, l4, l16))
    assert p4 < 0.03 and l4 < 0.01, (p4, l4)
    assert p16 < 0.5 * p4 and l16 < 0.5 * l4, (p4, p16, l4, l16)


def test_oracle_momentum_balance_in_free_flight():
    from oracle import oracle as O
    n = 64
    out = []
    for sub in (4, 16):
        cfg = O.default_config(K.ENV_IDS["AntMjBulletEnv-v0"], n)
        cfg.lin_damping = 0.0; cfg.ang_damping = 0.0; cfg.substeps = sub
        e = O.OracleVecEnv(cfg); e.reset()
        f, i = e.get_state()
        f = thrown_states(f, np.random.default_rng(0)); e.set_state(f, i)
        f0 = e.get_state()[0].copy()
        e.substeps(np.zeros((n, 8)), sub)
        out.append(residuals(f0, e.get_state()[0], range(n)))
    judge(out[0], out[1], "oracle")


@pytest.mark.gpu
def test_cuda_momentum_balance_in_free_flight_full_batch():
    import torch
    from hrl_pybullet_envs_b200 import VecEnv
    n = 4096
    out = []
    for sub in (4, 16):
        env = VecEnv("AntMjBulletEnv-v0", n, device=0, seed=0, config_overrides={"lin_damping": 0.0, "ang_damping": 0.0, "substeps": sub})
        env.reset()
        f, i = env.get_state()
        f0 = thrown_states(f.cpu().numpy().copy(), np.random.default_rng(1)).astype(np.float32)
        env.set_state(torch.tensor(f0).cuda(), i)
        f0 = env.get_state()[0].cpu().numpy()
        env.substeps(torch.zeros(n, 8, device="cuda"), sub)
        f1 = env.get_state()[0].cpu().numpy()
        assert np.isfinite(f1).all()
        out.append(residuals(f0, f1, range(0, n, 16)))
    judge(out[0], out[1], "cuda")
