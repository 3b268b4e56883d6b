"""Physics invariants of the CPU oracle (SURVEY.md section 4 tier iii), CPU only.

The oracle's physics is a restatement of Bullet's btMultiBody pipeline that cannot be compared with a
real Bullet here ("parity unpinned"); these tests hold it to what any correct
articulated-body step must satisfy: free fall, a symmetric positive-definite mass matrix consistent with
the model's total mass, joint limits, rest on the ground, determinism, and the reference's reset /
respawn placement rules.
"""
import numpy as np
import pytest

from hrl_pybullet_envs_b200 import config as K
from oracle import oracle as O

TOTAL_MASS = 65.44984694978736 + 8 * 7.831583314284915 + 4 * 13.51850726010076  # 1000 kg/m^3 x geom volumes = 182.1765
LO = np.array([-0.698132, 0.523599, -0.698132, -1.745329, -0.698132, -1.745329, -0.698132, 0.523599])
HI = np.array([0.698132, 1.745329, 0.698132, -0.523599, 0.698132, -0.523599, 0.698132, 1.745329])


def _env(env_id="AntMjBulletEnv-v0", n=4, **kw):
    e = O.OracleVecEnv.make(env_id, n, seed=1, **kw)
    e.reset()
    return e


def test_free_fall_acceleration():
    """No contacts, zero velocity, zero torque: every body accelerates with g, the joints do not move."""
    e = _env()
    f, i = e.get_state()
    f[:, K.SF_POS + 2] = 3.0  # lift the ants off the ground
    e.set_state(f, i)
    ud = e.free_accel(0)
    assert np.allclose(ud[:3], 0, atol=1e-9) and np.allclose(ud[3:6], [0, 0, -9.8], atol=1e-9)
    assert np.allclose(ud[6:], 0, atol=1e-8)


def test_inverse_mass_matrix_spd_and_total_mass():
    e = _env()
    Minv = e.inverse_mass_matrix(0)
    assert np.allclose(Minv, Minv.T, atol=1e-10)
    w = np.linalg.eigvalsh(Minv)
    assert w.min() > 0
    M = np.linalg.inv(Minv)
    # a uniform translation of the whole tree (joint velocities 0) has kinetic energy m v^2 / 2
    for ax in range(3):
        v = np.zeros(14); v[3 + ax] = 1.0
        assert v @ M @ v == pytest.approx(TOTAL_MASS, rel=1e-6)


def test_joint_torque_response_sign_and_magnitude():
    """A positive hip torque accelerates that hip positively and (reaction) turns the torso the other way."""
    e = _env()
    f, i = e.get_state(); f[:, K.SF_POS + 2] = 3.0; e.set_state(f, i)
    tau = np.zeros(8); tau[0] = 100.0
    f0 = e.free_accel(0)
    ud = e.free_accel(0, tau) - f0
    assert ud[6] > 0 and ud[2] < 0  # hip_1 about +z, torso yaw reaction
    assert 1.0 < ud[6] < 1e4


def test_drop_settles_on_the_ground_within_limits():
    """Zero action: the ant drops, stays alive and stands still; joints end inside their limits."""
    e = _env("AntMjBulletEnv-v0", 8)
    a = np.zeros((8, 8), np.float32)
    for t in range(200):
        obs, rew, done, info = e.step(a)
        assert np.isfinite(obs).all() and not done.any()
    f, _ = e.get_state()
    z = f[:, K.SF_POS + 2]
    assert (z > 0.26).all() and (z < 0.75).all()
    # it stands (on its feet, ankles on their 100 deg limit) and does not bounce; with only 5 solver iterations and
    # no joint damping (Bullet ignores the MJCF joint damping) the hips keep creeping slowly
    assert np.abs(f[:, K.SF_LINVEL:K.SF_LINVEL + 3]).max() < 0.2 and np.abs(f[:, K.SF_QD:K.SF_QD + 8]).max() < 1.0
    z0 = z.copy()
    for t in range(30):
        e.step(a)
    assert np.abs(e.get_state()[0][:, K.SF_POS + 2] - z0).max() < 5e-3
    q = f[:, K.SF_Q:K.SF_Q + 8]
    assert (q > LO - 0.02).all() and (q < HI + 0.02).all()  # ERP 0.2 pulls the ankles (reset at ~0) into [30, 100] deg
    quat = f[:, K.SF_QUAT:K.SF_QUAT + 4]
    assert np.allclose(np.linalg.norm(quat, axis=1), 1, atol=1e-9)


def test_random_actions_stay_finite_and_bounded():
    e = _env("AntGatherBulletEnv-v0", 32)
    rng = np.random.default_rng(0)
    for t in range(300):
        obs, rew, done, info = e.step(rng.uniform(-1, 1, (32, 8)).astype(np.float32))
        assert np.isfinite(obs).all()
        assert np.abs(obs[:, :26]).max() <= 5.0 + 1e-9  # WalkerBase.calc_state clips to +-5
    f, _ = e.get_state()
    assert np.abs(f[:, K.SF_POS:K.SF_POS + 2]).max() < 7.5  # inside the 15 x 15 arena walls
    assert np.abs(f[:, K.SF_LINVEL:K.SF_QD + 8]).max() <= 100.0 + 1e-9  # max_coord_vel clamp


def test_walls_contain_the_ant():
    """Flagrun arena (12 x 12): push the ant against a wall with a large initial velocity."""
    e = _env("AntFlagrunBulletEnv-v0", 2)
    f, i = e.get_state()
    f[:, K.SF_POS] = 5.0; f[:, K.SF_POS + 2] = 0.5; f[:, K.SF_LINVEL] = 8.0
    e.set_state(f, i)
    for t in range(60):
        e.step(np.zeros((2, 8), np.float32))
    f, _ = e.get_state()
    assert (f[:, K.SF_POS] < 5.95 + 0.05).all()  # inner wall face at 12/2 - 0.05, torso radius keeps the centre inside


def test_determinism_and_seed_sensitivity():
    a = np.random.default_rng(3).uniform(-1, 1, (50, 16, 8)).astype(np.float32)
    outs = []
    for seed in (1, 1, 2):
        e = O.OracleVecEnv.make("AntGatherBulletEnv-v0", 16, seed=seed)
        o0 = e.reset()
        for t in range(50):
            obs, rew, done, info = e.step(a[t])
        outs.append((o0, obs))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert not np.array_equal(outs[0][0], outs[2][0])


def test_reset_placement_rules():
    """gather_scene.py:38-62: items uniform on the (size - 1)^2 square, none within 2 m of the origin;
    WalkerBase.robot_specific_reset: joints ~ U(-0.1, 0.1), zero velocity, base at the MJCF pose."""
    e = O.OracleVecEnv.make("AntGatherBulletEnv-v0", 512, seed=7)
    e.reset()
    f, i = e.get_state()
    it = f[:, K.SF_ITEMS:K.SF_ITEMS + 32].reshape(512, 16, 2)
    assert np.abs(it).max() <= 7.0 and (np.linalg.norm(it, axis=2) >= 2.0).all()
    assert abs(it.mean()) < 0.2 and 3.5 < it.std() < 4.4          # U(-7, 7) with a hole: std ~ 4.1
    q = f[:, K.SF_Q:K.SF_Q + 8]
    assert np.abs(q).max() <= 0.1 and 0.05 < q.std() < 0.065       # U(-0.1, 0.1): std 0.0577
    assert (f[:, K.SF_QD:K.SF_QD + 8] == 0).all() and np.allclose(f[:, K.SF_POS:K.SF_POS + 3], [0, 0, 0.75])
    assert (i[:, K.SI_EPISODE] == 1).all() and (i[:, K.SI_T] == 0).all()


def test_pickup_respawn_rule():
    """ant_gather_env.py:86-92 + gather_scene.py:95-114: an item within 1 m (squared distance < 1) pays +-1 and
    respawns at least 2 m away from the robot."""
    e = O.OracleVecEnv.make("AntGatherBulletEnv-v0", 64, seed=5)
    e.reset()
    f, i = e.get_state()
    f[:, K.SF_ITEMS:K.SF_ITEMS + 2] = [0.5, 0.0]        # food 0 next to the torso
    f[:, K.SF_ITEMS + 16:K.SF_ITEMS + 18] = [0.0, -0.6]  # poison 8
    e.set_state(f, i)
    obs, rew, done, info = e.step(np.zeros((64, 8), np.float32))
    assert np.allclose(info[:, 0], 0.0) and np.allclose(rew, 0.0)  # +1 and -1 cancel
    f2, _ = e.get_state()
    for k in (0, 8):
        p = f2[:, K.SF_ITEMS + 2 * k:K.SF_ITEMS + 2 * k + 2]
        assert (np.linalg.norm(p - f2[:, K.SF_POS:K.SF_POS + 2], axis=1) >= 2.0 - 1e-3).all() and np.abs(p).max() <= 7.0


def test_capsule_cylinder_meets_the_maze_box_corner():
    """A foot capsule lying ACROSS the box corner (1, -2) (maze_scene.py:13: box x in [-5, 1], y in [-2, 2]) with both
    end-spheres clear of the box: only the cylinder-vs-vertical-edge test can see it.  The ant hangs at z = 1 (no
    ground contact); the corner contact must appear, push the leg away from the box, and vanish with has_box = 0."""
    def run(has_box):
        cfg = O.default_config(K.ENV_IDS["AntMazeBulletEnv-v0"], 1)
        cfg.has_box = has_box
        e = O.OracleVecEnv(cfg); e.reset()
        f, i = e.get_state()
        f[:] = 0
        # leg 1 (front-left, direction (+1, +1)): with q = 0 its foot capsule runs from O + 0.4 (1, 1) to O + 0.8 (1, 1).
        # Put its midpoint 0.07 m (< r = 0.08) outside the corner along the outward diagonal (+1, -1) / sqrt 2.
        mid = np.array([1.0, -2.0]) + 0.07 * np.array([1.0, -1.0]) / np.sqrt(2.0)
        f[0, K.SF_POS:K.SF_POS + 3] = [mid[0] - 0.6, mid[1] - 0.6, 1.0]
        f[0, K.SF_QUAT + 3] = 1.0
        f[0, K.SF_Q:K.SF_Q + 8] = [0, 1.6, 0, -1.6, 0, -1.6, 0, 1.6]   # other feet folded down (clear of box and ground), inside their ranges
        f[0, K.SF_Q + 1] = 0.0                                            # ... except leg 1: straight, so the geometry above holds
        e.set_state(f, i)
        e.substeps(np.zeros((1, 8), np.float32), 1)
        st = e.stats()
        return st, e.get_state()[0][0], f[0].copy()
    st1, f1, start = run(1)
    st0, f0, _ = run(0)
    # leg 1's straight ankle sits below its lower limit (30 deg): one limit row in both runs; the contact only with the box
    assert st0["contacts_per_substep"] == 0 and st1["contacts_per_substep"] == 1
    # the box contact is the only horizontal external force: the horizontal momentum of the whole tree (independent numpy
    # model of test_oracle_lagrangian) stays 0 without the box and is pushed along the outward normal (+1, -1) / sqrt 2 with it
    from test_oracle_lagrangian import links, _quat_R

    def momentum(f):
        # new velocities on the START configuration: the velocity update happens there (the position update that follows
        # moves the lever arms, which changes sum m J u at O(h) - generalized-coordinate Euler, Bullet does the same)
        u = np.concatenate([f[K.SF_ANGVEL:K.SF_ANGVEL + 3], f[K.SF_LINVEL:K.SF_LINVEL + 3], f[K.SF_QD:K.SF_QD + 8]])
        return sum(m * (Jv @ u) for m, _, _, _, Jv in links(start[K.SF_POS:K.SF_POS + 3], _quat_R(start[K.SF_QUAT:K.SF_QUAT + 4]),
                                                            start[K.SF_Q:K.SF_Q + 8]))
    p0, p1 = momentum(f0), momentum(f1)
    assert np.abs(p0[:2]).max() < 1e-3   # (link-constant tabulation differences of 4e-6 x a 25 rad/s limit correction)
    n = np.array([1.0, -1.0]) / np.sqrt(2.0)
    assert p1[:2] @ n > 0.1   # (the tangential part is friction: the foot swings under the limit correction)


def test_capsule_cylinder_vs_box_closest_point():
    """capsule_interior_vs_box (the cylinder part of a leg capsule against a food / poison cube): the bisection on the
    derivative of the squared point-box distance finds the segment's interior point closest to the box - checked against
    a brute-force scan - and reports nothing when the closest point is an END of the segment (the end spheres' job)."""
    import ctypes as C
    L = O.lib()
    L.hrlo_capsule_vs_box.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    lo = np.array([-0.125, -0.125, -0.025], np.float32); hi = np.array([0.125, 0.125, 0.225], np.float32)
    found = ends = 0
    for _ in range(400):
        A = rng.uniform(-0.6, 0.6, 3); B = A + rng.normal(size=3) * 0.4
        A[2] = abs(A[2]); B[2] = abs(B[2])
        out = np.zeros(7)
        ok = L.hrlo_capsule_vs_box(O._p(A), O._p(B), 0.08, O._p(lo), O._p(hi), O._p(out))
        t = np.linspace(0, 1, 20001)[:, None]
        P = A + t * (B - A)
        e = np.maximum(P - hi, 0) + np.minimum(P - lo, 0)
        d = np.linalg.norm(e, axis=1)
        j = int(d.argmin())
        interior = 0 < j < 20000 and d[j] > 0
        if ok:
            found += 1
            if d.min() > 0:
                assert abs(out[6] - (d.min() - 0.08)) < 1e-5, (out[6], d.min())
                np.testing.assert_allclose(out[:3], P[j], atol=2e-4)
                np.testing.assert_allclose(out[3:6], e[j] / d[j], atol=2e-3)
        else:
            ends += 1
            assert not interior or d[j] < 1e-9 or min(j, 20000 - j) < 3, (j, d[j])   # minimum at an end (or the segment starts inside)
    assert found > 50 and ends > 50
