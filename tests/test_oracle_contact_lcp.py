"""The oracle's constraint solve against the INDEPENDENT numpy model of test_oracle_lagrangian (CPU only).

With many solver iterations the projected Gauss-Seidel sweep converges to the solution of the contact / joint-limit
complementarity problem it sets up.  This test measures that solution with something other than the oracle's own
Jacobians: contact-point velocities from the naive maximal-coordinate kinematics (links() of test_oracle_lagrangian)
on the START configuration with the END-of-sub-step velocities, and checks Bullet's row conventions as restated in
SURVEY.md App. A.3 step 4:

  * every contact (sphere within the 0.02 margin of the ground): normal velocity v_n >= v*, with
    v* = -dist * erp / h for a penetrating contact (dist <= 0, erp = 0.9) and v* = -dist / h for a separated one;
  * the states are taken while the ant lands from its 0.3 m reset drop and then stands under random actions, so the
    rows stop real impacts (normal velocities of 1-2 m/s before the solve);
  * every joint at or beyond a limit: joint velocity >= -pen * 0.2 / h towards the inside.

A sign error in a Jacobian row, a wrong lever arm or link assignment of a contact point, or a wrong right-hand side
shows up here at O(1) m/s against a tolerance of 5e-3.
"""
import numpy as np

from hrl_pybullet_envs_b200 import config as K
from oracle import oracle as O
from test_oracle_lagrangian import SIGNS, ANKLE_AXIS, R_CAPS, R_TORSO, _quat_R, _rot, links

H = 0.0165 / 4


def _spheres(pos, R, q):
    """(link index in links(), world centre, radius) of the 13 collision spheres: torso, per leg tip / ankle / hip."""
    out = [(0, pos.copy(), R_TORSO)]
    for k, (sx, sy) in enumerate(SIGNS):
        d = np.array([sx, sy, 0.0])
        hip = pos + R @ (0.2 * d)
        Ra = R @ _rot((0, 0, 1), q[2 * k])
        ank = hip + Ra @ (0.2 * d)
        Rf = Ra @ _rot(np.asarray(ANKLE_AXIS[k], float), q[2 * k + 1])
        tip = ank + Rf @ (0.4 * d)
        out += [(3 + 3 * k, tip, R_CAPS), (2 + 3 * k, ank, R_CAPS), (1 + 3 * k, hip, R_CAPS)]
    return out


def _u(f):
    return np.concatenate([f[K.SF_ANGVEL:K.SF_ANGVEL + 3], f[K.SF_LINVEL:K.SF_LINVEL + 3], f[K.SF_QD:K.SF_QD + 8]])


TOL = 1e-4  # m/s after 400 iterations; impact velocities below are ~2 m/s.  The converged sub-step is run WITHOUT friction:
#             normal and limit rows alone form a linear complementarity problem with a positive semi-definite matrix, on
#             which the sweep converges to the solution.  With Bullet's implicit friction cone (the two rows of a pair are
#             updated from the same velocities, then scaled into the cone) it can settle into a 2-cycle between the
#             alternating row orders instead - measured here on a foot pushed sideways by a hip torque at its limit - so
#             the end of a sweep is not a solution of every row; that is a property of the algorithm, not of the rows.


def test_converged_sweep_solves_the_contact_and_limit_lcp():
    n = 8
    cfg = O.default_config(K.ENV_IDS["AntMjBulletEnv-v0"], n)   # flat ground at z = 0, no walls; reset drops the ant 0.3 m
    e = O.OracleVecEnv(cfg); e.reset()
    cfg2 = cfg.copy(); cfg2.solver_iters = 400; cfg2.friction = 0.0
    e2 = O.OracleVecEnv(cfg2); e2.reset()
    rng = np.random.default_rng(4)
    lo = np.array([-0.698132, 0.523599, -0.698132, -1.745329, -0.698132, -1.745329, -0.698132, 0.523599])
    hi = np.array([0.698132, 1.745329, 0.698132, -0.523599, 0.698132, -0.523599, 0.698132, 1.745329])
    n_contacts = n_impacts = n_active = n_limits = 0
    for t in range(60):
        act = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        if t >= 10:  # from shortly before the landing on: one converged sub-step from the current state, measured independently
            f0, i0 = e.get_state()
            e2.set_state(f0, i0); e2.substeps(act, 1)
            f1, _ = e2.get_state()
            for j in range(n):
                pos, R, q = f0[j, K.SF_POS:K.SF_POS + 3], _quat_R(f0[j, K.SF_QUAT:K.SF_QUAT + 4]), f0[j, K.SF_Q:K.SF_Q + 8]
                u0, u1 = _u(f0[j]), _u(f1[j])
                L = links(pos, R, q)
                per_group = [0, 0, 0, 0]
                for si, (li, c, r) in enumerate(_spheres(pos, R, q)):
                    dist = c[2] - r - cfg.ground_z
                    if not dist < cfg.contact_margin:
                        continue
                    g = 0 if si == 0 else (si - 1) // 3
                    if per_group[g] >= 4:
                        continue
                    per_group[g] += 1
                    m, I, com, Jw, Jv = L[li]
                    P = c - np.array([0, 0, r])
                    v_before = (Jv @ u0 + np.cross(Jw @ u0, P - com))[2]
                    vP = Jv @ u1 + np.cross(Jw @ u1, P - com)
                    vstar = -dist * cfg.contact_erp / H if dist <= 0 else -dist / H
                    n_contacts += 1
                    n_impacts += v_before < vstar - 0.3
                    assert vP[2] >= vstar - TOL, (t, j, si, v_before, vP[2], vstar)
                    n_active += abs(vP[2] - vstar) < TOL
                for d in range(8):
                    pl, ph = q[d] - lo[d], hi[d] - q[d]
                    if pl <= 0:
                        n_limits += 1
                        assert u1[6 + d] >= -pl * cfg.limit_erp / H - TOL, (t, j, d)
                    if ph <= 0:
                        n_limits += 1
                        assert -u1[6 + d] >= -ph * cfg.limit_erp / H - TOL, (t, j, d)
        e.step(act)
    print("contacts %d (impacting at > 0.3 m/s: %d, loaded after the solve: %d), joint-limit rows %d" % (n_contacts, n_impacts, n_active, n_limits))
    assert n_contacts > 500 and n_impacts >= 8 and n_active > 300 and n_limits > 20


def test_friction_opposes_sliding_inside_the_cone():
    """Ants standing on the ground are given a horizontal velocity: the horizontal momentum (numpy model) lost in one
    sub-step is the total friction impulse, the vertical momentum gained over free fall is the total normal impulse.
    Friction must oppose the sliding direction with a magnitude of the order of mu N (mu = 1.5 x 0.8)."""
    n = 8
    cfg = O.default_config(K.ENV_IDS["AntMjBulletEnv-v0"], n)
    e = O.OracleVecEnv(cfg); e.reset()
    zero = np.zeros((n, 8), np.float32)
    for t in range(120):
        e.step(zero)                                            # settle on the feet
    f0, i0 = e.get_state()
    rng = np.random.default_rng(2)
    ang = rng.uniform(0, 2 * np.pi, n)
    slide = np.stack([np.cos(ang), np.sin(ang)], 1) * rng.uniform(0.5, 3.0, (n, 1))
    f0[:, K.SF_LINVEL:K.SF_LINVEL + 2] = slide
    e.set_state(f0, i0)
    e.substeps(zero, 1)
    f1, _ = e.get_state()
    mass = sum(m for m, _, _, _, _ in links(f0[0, 0:3], _quat_R(f0[0, 3:7]), f0[0, 13:21]))
    for j in range(n):
        L = links(f0[j, K.SF_POS:K.SF_POS + 3], _quat_R(f0[j, K.SF_QUAT:K.SF_QUAT + 4]), f0[j, K.SF_Q:K.SF_Q + 8])
        p0 = sum(m * (Jv @ _u(f0[j])) for m, _, _, _, Jv in L)
        p1 = sum(m * (Jv @ _u(f1[j])) for m, _, _, _, Jv in L)
        dp = p1 - p0
        normal = dp[2] + mass * 9.8 * H                          # what the ground added on top of gravity
        fric = dp[:2]
        s = slide[j] / np.linalg.norm(slide[j])
        assert normal > 0.5 * mass * 9.8 * H                     # the ground carries the ant
        assert fric @ s < 0                                       # friction brakes the slide ...
        # (not collinear with it: the impulse that stops a foot of an articulated body is M^-1-weighted, and the cone only
        # scales the unconstrained pair solution radially)
        # Magnitude: of the order of mu N.  Not a strict cone at the level of the whole ant: a friction pair whose normal
        # impulse has dropped back to zero during the sweep is skipped with its impulses KEPT (Bullet's `totalImpulse > 0`
        # guard [3P-MEM], restated in the oracle and the kernel), so the sum can exceed mu x sum(N) - measured up to
        # 1.37 mu N after 5 iterations on these poses.  (Air drag, Bullet's 0.04 damping, is < 1 % of these impulses.)
        assert 0.4 * cfg.friction * normal < np.linalg.norm(fric) < 1.5 * cfg.friction * normal, (j, np.linalg.norm(fric), cfg.friction * normal)
