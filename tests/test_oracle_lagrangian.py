"""The oracle's articulated-body dynamics against an INDEPENDENT maximal-coordinate model (CPU only).

The oracle (oracle/hrl_oracle.c) computes the unconstrained step with Featherstone's ABA in link coordinates;
the CUDA path uses a world-frame projected Newton-Euler with a block-arrow mass matrix.  Neither can be compared
with a real Bullet here ("parity unpinned").  This file adds a third, deliberately naive
formulation built straight from the reference's model file (hrl_pybullet_envs/assets/ant.xml:12-56: torso sphere,
per leg a rigid capsule, a hip hinge about z and an ankle hinge about (-+1, 1, 0)) in numpy:

  * link poses / velocities by explicit forward kinematics, Jacobians column by column from rigid-body kinematics;
  * the mass matrix as  M = sum_i  m_i Jv_i^T Jv_i + Jw_i^T I_i Jw_i;
  * Kane's equations  sum_i [ m_i a_i . Jv_i e_k + (I_i alpha_i + w_i x I_i w_i) . Jw_i e_k ] = generalized force_k,
    with the link accelerations a_i, alpha_i taken by central differences ALONG the trajectory that the oracle's own
    generalized acceleration defines.

Agreement of all 14 components at random poses, joint angles and velocities pins the oracle's kinematic tree, its
inertial parameters (Bullet's rules as recalled in SURVEY.md App. A.3: 1000 kg/m^3 x geom volume, inertia of the
AABB box in link axes) and its Coriolis / centrifugal / gyroscopic terms on something other than itself.
"""
import numpy as np
import pytest

from hrl_pybullet_envs_b200 import config as K
from oracle import oracle as O

SIGNS = [(1, 1), (-1, 1), (-1, -1), (1, -1)]                 # ant.xml:15,26,37,48 (front_left, front_right, back, right_back)
ANKLE_AXIS = [(-1, 1, 0), (1, 1, 0), (-1, 1, 0), (1, 1, 0)]  # ant.xml:21,32,43,54
R_CAPS, R_TORSO, RHO = 0.08, 0.25, 1000.0                     # ant.xml:13,16; Bullet's MJCF importer density [3P-MEM]
G = np.array([0.0, 0.0, -9.8])


def _capsule(length):
    """mass and (ixx = iyy, izz) of a capsule whose axis is a horizontal diagonal: Bullet's AABB-box rule in link axes."""
    m = RHO * (np.pi * R_CAPS ** 2 * length + 4.0 / 3.0 * np.pi * R_CAPS ** 3)
    lx = ly = length / np.sqrt(2.0) + 2 * R_CAPS
    lz = 2 * R_CAPS
    return m, m / 12.0 * (ly * ly + lz * lz), m / 12.0 * (lx * lx + ly * ly)


M_TORSO = RHO * 4.0 / 3.0 * np.pi * R_TORSO ** 3
I_TORSO = M_TORSO / 12.0 * 2 * (2 * R_TORSO) ** 2
M_S, IX_S, IZ_S = _capsule(0.2 * np.sqrt(2.0))
M_L, IX_L, IZ_L = _capsule(0.4 * np.sqrt(2.0))


def _quat_R(q):  # xyzw
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _rot(axis, ang):
    a = np.asarray(axis, float) / np.linalg.norm(axis)
    Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def links(pos, R, q):
    """13 links: (mass, world inertia tensor, world COM, Jw[3x14], Jv[3x14]) in the oracle's generalized velocity
    [w_world(3), v_world of the torso origin(3), qdot(8)], joints ordered hip_1, ankle_1, ..., hip_4, ankle_4."""
    out = []

    def base_cols(r):  # velocity of a torso-fixed point at world offset r from the torso origin
        Jw = np.zeros((3, 14)); Jv = np.zeros((3, 14))
        Jw[:, 0:3] = np.eye(3); Jv[:, 3:6] = np.eye(3); Jv[:, 0:3] = -_skew(r)
        return Jw, Jv
    Jw, Jv = base_cols(np.zeros(3))
    out.append((M_TORSO, I_TORSO * np.eye(3), pos.copy(), Jw, Jv))
    for k, (sx, sy) in enumerate(SIGNS):
        d = np.array([sx, sy, 0.0])
        # rigid leg capsule on the torso (ant.xml:15-16)
        c = R @ (0.1 * d)
        Jw, Jv = base_cols(c)
        I = R @ np.diag([IX_S, IX_S, IZ_S]) @ R.T
        out.append((M_S, I, pos + c, Jw, Jv))
        # aux link behind the hip hinge (axis z at 0.2 d, ant.xml:17-19)
        hip = R @ (0.2 * d)
        Ra = R @ _rot((0, 0, 1), q[2 * k])
        a1 = R @ np.array([0.0, 0.0, 1.0])
        c = hip + Ra @ (0.1 * d)
        Jw, Jv = base_cols(c)
        Jw[:, 6 + 2 * k] = a1; Jv[:, 6 + 2 * k] = np.cross(a1, c - hip)
        out.append((M_S, Ra @ np.diag([IX_S, IX_S, IZ_S]) @ Ra.T, pos + c, Jw, Jv))
        # foot link behind the ankle hinge (at 0.2 d in the aux frame, ant.xml:20-23)
        ank = hip + Ra @ (0.2 * d)
        ax = np.asarray(ANKLE_AXIS[k], float) / np.sqrt(2.0)
        Rf = Ra @ _rot(ax, q[2 * k + 1])
        a2 = Ra @ ax
        c = ank + Rf @ (0.2 * d)
        Jw, Jv = base_cols(c)
        Jw[:, 6 + 2 * k] = a1; Jv[:, 6 + 2 * k] = np.cross(a1, c - hip)
        Jw[:, 7 + 2 * k] = a2; Jv[:, 7 + 2 * k] = np.cross(a2, c - ank)
        out.append((M_L, Rf @ np.diag([IX_L, IX_L, IZ_L]) @ Rf.T, pos + c, Jw, Jv))
    return out


def mass_matrix(pos, R, q):
    M = np.zeros((14, 14))
    for m, I, _, Jw, Jv in links(pos, R, q):
        M += m * Jv.T @ Jv + Jw.T @ I @ Jw
    return M


def _random_state(rng, e, n):
    f, i = e.get_state()
    for j in range(n):
        qt = rng.normal(size=4); qt /= np.linalg.norm(qt)
        f[j, K.SF_POS:K.SF_POS + 3] = [rng.uniform(-1, 1), rng.uniform(-1, 1), 3.0]
        f[j, K.SF_QUAT:K.SF_QUAT + 4] = qt
        lo = np.array([-0.6, 0.6, -0.6, -1.6, -0.6, -1.6, -0.6, 0.6]); hi = np.array([0.6, 1.6, 0.6, -0.6, 0.6, -0.6, 0.6, 1.6])
        f[j, K.SF_Q:K.SF_Q + 8] = rng.uniform(lo, hi)
        f[j, K.SF_LINVEL:K.SF_LINVEL + 3] = rng.normal(size=3)
        f[j, K.SF_ANGVEL:K.SF_ANGVEL + 3] = 2 * rng.normal(size=3)
        f[j, K.SF_QD:K.SF_QD + 8] = 3 * rng.normal(size=8)
    e.set_state(f, i)
    return f


def _env(n, damping):
    cfg = O.default_config(K.ENV_IDS["AntMjBulletEnv-v0"], n)
    if not damping:
        cfg.lin_damping = 0.0; cfg.ang_damping = 0.0
    e = O.OracleVecEnv(cfg)
    e.reset()
    return e


def test_model_constants():
    assert M_TORSO + 8 * M_S + 4 * M_L == pytest.approx(182.1765, abs=1e-3)


def test_mass_matrix_matches_independent_model():
    n = 6
    e = _env(n, damping=False)
    f = _random_state(np.random.default_rng(3), e, n)
    for j in range(n):
        M = mass_matrix(f[j, K.SF_POS:K.SF_POS + 3], _quat_R(f[j, K.SF_QUAT:K.SF_QUAT + 4]), f[j, K.SF_Q:K.SF_Q + 8])
        Minv = e.inverse_mass_matrix(j)
        # 5e-5: the oracle's tabulated link masses / inertias differ from the closed formulas above
        # by 4e-6 relative; a wrong axis, offset or parent would show up at O(0.1)
        assert np.allclose(Minv @ M, np.eye(14), atol=5e-5), np.abs(Minv @ M - np.eye(14)).max()


def test_free_acceleration_satisfies_kanes_equations():
    """The oracle's generalized acceleration (gravity + joint torques, Bullet's damping switched off) closes Kane's
    equations of the independent model in all 14 directions: mass matrix AND velocity-dependent bias terms."""
    n = 6
    rng = np.random.default_rng(11)
    e = _env(n, damping=False)
    f = _random_state(rng, e, n)
    h = 1e-5
    for j in range(n):
        tau = rng.uniform(-100, 100, 8)
        ud = e.free_accel(j, tau)
        pos = f[j, K.SF_POS:K.SF_POS + 3]; R = _quat_R(f[j, K.SF_QUAT:K.SF_QUAT + 4]); q = f[j, K.SF_Q:K.SF_Q + 8]
        u = np.concatenate([f[j, K.SF_ANGVEL:K.SF_ANGVEL + 3], f[j, K.SF_LINVEL:K.SF_LINVEL + 3], f[j, K.SF_QD:K.SF_QD + 8]])

        def at(dt):  # configuration and velocity a time dt along the trajectory (second order in dt)
            w = u[0:3] * dt + 0.5 * ud[0:3] * dt * dt
            ang = np.linalg.norm(w)
            Rt = (_rot(w, ang) if ang > 0 else np.eye(3)) @ R
            return pos + u[3:6] * dt + 0.5 * ud[3:6] * dt * dt, Rt, q + u[6:] * dt + 0.5 * ud[6:] * dt * dt, u + ud * dt
        Lm, Lp, L0 = links(*at(-h)[:3]), links(*at(h)[:3]), links(pos, R, q)
        um, up = at(-h)[3], at(h)[3]
        resid = np.zeros(14); scale = np.zeros(14)
        for (m, I, _, Jw, Jv), lm, lp in zip(L0, Lm, Lp):
            a = (lp[4] @ up - lm[4] @ um) / (2 * h)            # COM acceleration
            Lang = (lp[1] @ (lp[3] @ up) - lm[1] @ (lm[3] @ um)) / (2 * h)  # d/dt (I w) = I alpha + w x I w
            resid += Jv.T @ (m * a - m * G) + Jw.T @ Lang
            scale += np.abs(Jv.T @ (m * a)) + np.abs(Jw.T @ Lang) + np.abs(Jv.T @ (m * G))
        resid[6:] -= tau
        scale[6:] += np.abs(tau)
        assert np.all(np.abs(resid) < 3e-5 * (1 + scale)), (resid, scale)


def test_damping_is_bullets_per_link_drag():
    """With the default damping on, the extra generalized force equals -sum_i [k m_i (1 + |v_i|) v_i . Jv_i +
    k (1 + |w_i|) (I_i w_i) . Jw_i], k = 0.04: Bullet's per-link linear / angular damping [3P-MEM]."""
    n = 4
    rng = np.random.default_rng(5)
    e0, e1 = _env(n, damping=False), _env(n, damping=True)
    f = _random_state(rng, e0, n)
    e1.set_state(f, e1.get_state()[1])
    for j in range(n):
        pos = f[j, K.SF_POS:K.SF_POS + 3]; R = _quat_R(f[j, K.SF_QUAT:K.SF_QUAT + 4]); q = f[j, K.SF_Q:K.SF_Q + 8]
        u = np.concatenate([f[j, K.SF_ANGVEL:K.SF_ANGVEL + 3], f[j, K.SF_LINVEL:K.SF_LINVEL + 3], f[j, K.SF_QD:K.SF_QD + 8]])
        Q = np.zeros(14)
        for m, I, _, Jw, Jv in links(pos, R, q):
            v, w = Jv @ u, Jw @ u
            Q -= Jv.T @ (0.04 * m * (1 + np.linalg.norm(v)) * v) + Jw.T @ (0.04 * (1 + np.linalg.norm(w)) * (I @ w))
        d_ud = e1.free_accel(j) - e0.free_accel(j)
        assert np.allclose(mass_matrix(pos, R, q) @ d_ud, Q, rtol=1e-4, atol=1e-4)
