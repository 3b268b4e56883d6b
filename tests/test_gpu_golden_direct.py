"""The CUDA path against the reference's STEP-LEVEL golden vectors, directly (no oracle in between).

tests/golden/{gather_step, maze_step, maze_mj_step, flagrun_step, robots}.npz hold outputs of the reference's own
`step()` methods executed with a stub robot (tests/golden/make_golden.py).  Here the fixture's robot pose is written into
the kernels' state with `hrl_set_state` and ONE control step is taken with `substeps = 0` - the physics loop runs zero
times, so the kernel's task layer (pickups, sensors, goal logic, reward, done, counters) works on exactly the pose the
reference saw.  What the stub robot made up (the 28-d `calc_state` vector, the scripted `walk_target_dist`) cannot be
injected - the kernels compute those from the physical state - so each test compares the quantities that are a function
of the injected pose only, and says which.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from hrl_pybullet_envs_b200 import config as K  # noqa: E402

pytestmark = pytest.mark.gpu


def _quat_from_rpy(rpy):
    """Bullet getQuaternionFromEuler (x, y, z, w) for roll, pitch, yaw."""
    r, p, y = rpy[:, 0] / 2, rpy[:, 1] / 2, rpy[:, 2] / 2
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy], 1)


def _env(env_id, n, **kw):
    from hrl_pybullet_envs_b200 import VecEnv
    over = dict(substeps=0)
    over.update(kw.pop("config_overrides", {}))
    e = VecEnv(env_id, n, seed=1, auto_reset=False, config_overrides=over, **kw)
    e.reset()
    return e


def _set(env, f, i):
    env.set_state(torch.tensor(f, dtype=torch.float32), torch.tensor(i, dtype=torch.int32))


# ------------------------------------------------------------------ AntGatherBulletEnv.step (gather_step.npz)
@pytest.mark.parametrize("tag,kw", [("ant", {}), ("antabs", dict(use_sensor=False, n_bins=5))])
def test_gather_step_golden_on_gpu(golden, tag, kw):
    """Pickups (count and sign), alive / done, reward = food_rew + dead_rew for every fixture row; the sensor part of the
    observation (sector readings, or the nearest-item coordinates of use_sensor=False) for the rows in which no item
    was picked up (a respawn draws from Philox here and from the replayed MT19937 uniforms there)."""
    g = golden("gather_step.npz")
    xyz, rpy, objs = g[f"{tag}_xyz"], g[f"{tag}_rpy"], g[f"{tag}_objs"]
    M = len(xyz)
    env = _env("AntGatherBulletEnv-v0", M, item_contacts=False, **kw)
    f, i = env.get_state(); f = f.cpu().numpy().astype(np.float64); i = i.cpu().numpy()
    f[:, K.SF_POS:K.SF_POS + 3] = xyz
    f[:, K.SF_QUAT:K.SF_QUAT + 4] = _quat_from_rpy(rpy)
    f[:, K.SF_LINVEL:K.SF_LINVEL + 6] = 0; f[:, K.SF_Q:K.SF_Q + 16] = 0
    f[:, K.SF_INITIAL_Z] = 0.75
    f[:, K.SF_ITEMS:K.SF_ITEMS + 32] = objs.reshape(M, 32)
    _set(env, f, i)
    obs, rew, done, info = env.step(torch.zeros(M, 8, device="cuda"))
    obs = obs.cpu().numpy(); rew = rew.cpu().numpy(); done = done.cpu().numpy()
    ok = np.isfinite(g[f"{tag}_state"]).all(axis=1)           # rows 5, 6 carry the reference's non-finite guard (stub state)
    np.testing.assert_allclose(info["food_rew"].cpu().numpy()[ok], g[f"{tag}_food_rew"][ok], atol=0)
    np.testing.assert_allclose(info["dead_rew"].cpu().numpy()[ok], g[f"{tag}_dead_rew"][ok], atol=0)
    np.testing.assert_allclose(rew[ok], g[f"{tag}_rew"][ok], atol=0)
    assert np.array_equal(done[ok], g[f"{tag}_done"][ok])
    assert np.abs(obs[ok, 0] - g[f"{tag}_obs"][ok, 0]).max() < 1e-6      # z - initial_z
    still = ok & (np.abs(g[f"{tag}_new_objs"] - objs).max(axis=(1, 2)) == 0)
    assert still.sum() > 100 and (~still & ok).sum() > 50                  # both kinds of rows are in the fixture
    want = g[f"{tag}_obs"][still, 26:]
    got = obs[still, 26:]
    assert want.shape == got.shape
    if tag == "ant":      # bins bit-exact, intensities to 1e-5 (north star)
        assert np.array_equal(got != 0, want != 0)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-5)
    f2, _ = env.get_state()
    items = f2.cpu().numpy()[:, K.SF_ITEMS:K.SF_ITEMS + 32].reshape(M, 16, 2)
    moved_ref = np.abs(g[f"{tag}_new_objs"] - objs).max(axis=2) > 0
    moved_gpu = np.abs(items - objs.astype(np.float32)).max(axis=2) > 0
    assert np.array_equal(moved_gpu[ok], moved_ref[ok])                    # the same items were respawned


# ------------------------------------------------------------------ AntMazeBulletEnv.step (maze_step.npz)
@pytest.mark.parametrize("tag,kw", [("", {}), ("_angle_nowalls", dict(target_encoding=1, sense_walls=False))])
def test_maze_step_obs_golden_on_gpu(golden, tag, kw):
    """Goal part (vector or angle encoding, from the TRUE torso xy) and the wall lidar of the observation."""
    g = golden("maze_step.npz")
    M = len(g["xy"])
    env = _env("AntMazeBulletEnv-v0", M, **kw)
    f, i = env.get_state(); f = f.cpu().numpy().astype(np.float64); i = i.cpu().numpy()
    f[:, K.SF_POS] = g["xy"][:, 0]; f[:, K.SF_POS + 1] = g["xy"][:, 1]; f[:, K.SF_POS + 2] = 0.5
    rpy = np.zeros((M, 3)); rpy[:, 2] = g["yaw"]
    f[:, K.SF_QUAT:K.SF_QUAT + 4] = _quat_from_rpy(rpy)
    f[:, K.SF_TARGET:K.SF_TARGET + 2] = g["targets"][g["tid"]]
    _set(env, f, i)
    obs, rew, done, info = env.step(torch.zeros(M, 8, device="cuda"))
    want = g["obs" + tag]
    assert obs.shape == want.shape
    np.testing.assert_allclose(obs.cpu().numpy()[:, 26:], want[:, 26:], rtol=0, atol=2e-6)


def test_maze_mj_step_obs_golden_on_gpu(golden):
    """AntMazeMjEnv: MjAnt pass-through columns, lidar cast from the observation's own xy, the two zero blocks and the
    step counter t * 0.001 taken BEFORE the increment (ant_maze_mj_env.py:57-71)."""
    g = golden("maze_mj_step.npz")
    mj = g["mj_obs"]; M = len(mj)
    env = _env("AntMazeMjEnv-v0", M)
    f, i = env.get_state(); f = f.cpu().numpy().astype(np.float64); i = i.cpu().numpy()
    f[:, K.SF_POS:K.SF_POS + 3] = mj[:, 0:3]
    rpy = np.zeros((M, 3)); rpy[:, 2] = g["yaw"]
    f[:, K.SF_QUAT:K.SF_QUAT + 4] = _quat_from_rpy(rpy)        # (the fixture's quaternion columns are not unit: not injected)
    f[:, K.SF_Q:K.SF_Q + 8] = mj[:, 7:15]; f[:, K.SF_LINVEL:K.SF_LINVEL + 3] = mj[:, 15:18]
    f[:, K.SF_ANGVEL:K.SF_ANGVEL + 3] = mj[:, 18:21]; f[:, K.SF_QD:K.SF_QD + 8] = mj[:, 21:29]
    i[:, K.SI_T] = g["t_before"]
    _set(env, f, i)
    obs, rew, done, info = env.step(torch.zeros(M, 8, device="cuda"))
    obs = obs.cpu().numpy(); want = g["obs"]
    cols = [c for c in range(60) if not 3 <= c < 7]
    np.testing.assert_allclose(obs[:, cols], want[:, cols], rtol=0, atol=2e-6)
    _, i2 = env.get_state()
    assert np.array_equal(i2.cpu().numpy()[:, K.SI_T], g["t_after"])


# ------------------------------------------------------------------ AntFlagrunBulletEnv.step (flagrun_step.npz)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_flagrun_step_sequence_golden_on_gpu(golden, tag):
    """The scripted walk_target_dist sequence of the fixture is realised geometrically: before every step the ant is
    put (symmetric pose, so that the mean over its 13 parts is the torso) where the kernel's own quirk-Q1 distance to
    its CURRENT goal equals the scripted value.  Compared per step: the goal bonus (reward minus the inner reward),
    done, steps_since_goal_change, the _rewarded flag and whether the goal switched (ant_flagrun_env.py:162-204)."""
    g = golden("flagrun_step.npz")
    wtd = g["wtd"]; T = len(g[f"{tag}_rew"])
    env = _env("AntFlagrunBulletEnv-v0", 1, max_targets=int(g[f"{tag}_n_goals"]), timeout=int(g[f"{tag}_timeout"]),
               switch_flag_on_collision=bool(g[f"{tag}_switch"]))
    f, i = env.get_state()
    assert int(i[0, K.SI_GOALS_LEFT]) == int(g[f"{tag}_n_goals"]) - 1     # reset popped the first goal
    prev_target = g[f"{tag}_first_target"]
    for t in range(T):
        f, i = env.get_state(); f = f.cpu().numpy().astype(np.float64); i = i.cpu().numpy()
        tgt = f[0, K.SF_TARGET:K.SF_TARGET + 2].copy()
        # body_xy = (13 O + (-6, 0)) / 15 for a point-symmetric pose (quirk Q1: floor and wall averaged in)
        d = np.array([np.cos(0.7 * t), np.sin(0.7 * t)])
        body = tgt - wtd[t] * d
        O = (15 * body + np.array([6.0, 0.0])) / 13
        f[0, K.SF_POS:K.SF_POS + 3] = [O[0], O[1], 0.5]
        f[0, K.SF_QUAT:K.SF_QUAT + 4] = [0, 0, 0, 1]
        f[0, K.SF_Q:K.SF_Q + 8] = [0, 1.0, 0, -1.0, 0, -1.0, 0, 1.0]; f[0, K.SF_QD:K.SF_QD + 8] = 0
        f[0, K.SF_LINVEL:K.SF_LINVEL + 6] = 0
        _set(env, f, i)
        env.observe()
        got_wtd = float(env.get_state()[0][0, K.SF_WTD])
        assert abs(got_wtd - wtd[t]) < 2e-5, (t, got_wtd, wtd[t])
        obs, rew, done, info = env.step(torch.zeros(1, 8, device="cuda"))
        bonus = float(rew[0]) - float(info["inner_rew"][0])
        want_bonus = g[f"{tag}_rew"][t] - g["inner_r"][t]
        assert abs(bonus - want_bonus) < 0.51, (t, bonus, want_bonus)      # 0 or +5000 (f32 rounding of 5000 + inner)
        f2, i2 = env.get_state(); i2 = i2.cpu().numpy()
        assert bool(done[0]) == bool(g[f"{tag}_done"][t]), t
        if not done[0]:
            assert int(i2[0, K.SI_SINCE]) == int(g[f"{tag}_since"][t]), t
            assert bool(i2[0, K.SI_REWARDED]) == bool(g[f"{tag}_rewarded"][t]), t
        switched_ref = not np.array_equal(g[f"{tag}_target"][t], prev_target)
        prev_target = g[f"{tag}_target"][t]
        assert bool(info["target_switched"][0]) == switched_ref, t
        if done[0]:
            break
    assert t == T - 1


# ------------------------------------------------------------------ AntMjEnv.step reward composition (robots.npz)
def test_mj_reward_golden_on_gpu(golden):
    """alive(z) + progress + joints_at_limit_cost * #limits and done (envs/MjAnt.py:36-97): z, the old potential and the
    joints at their limits are injected; the new potential follows from the kernel's own walk_target_dist, so the old
    one is shifted to reproduce the fixture's progress term."""
    g = golden("robots.npz")
    st, jal = g["mj_state"], g["jal"]; M = len(st)
    env = _env("AntMjBulletEnv-v0", M)
    f, i = env.get_state(); f = f.cpu().numpy().astype(np.float64); i = i.cpu().numpy()
    f[:, K.SF_POS] = 0.3; f[:, K.SF_POS + 1] = 0.0; f[:, K.SF_POS + 2] = st[:, 2]
    f[:, K.SF_QUAT:K.SF_QUAT + 4] = [0, 0, 0, 1]; f[:, K.SF_LINVEL:K.SF_LINVEL + 6] = 0; f[:, K.SF_QD:K.SF_QD + 8] = 0
    f[:, K.SF_TARGET:K.SF_TARGET + 2] = 0.0                                # walk target at the origin: distances stay O(1)
    lim = np.array([0.697, 1.744, 0.697, -1.744, 0.697, -1.744, 0.697, 1.744])      # |rel| > 0.99
    mid = np.array([0.0, 1.13, 0.0, -1.13, 0.0, -1.13, 0.0, 1.13])
    q = np.tile(mid, (M, 1))
    for m in range(M):
        q[m, :jal[m]] = lim[:jal[m]]
    f[:, K.SF_Q:K.SF_Q + 8] = q
    _set(env, f, i)
    env.observe()
    f1, i1 = env.get_state(); f1 = f1.cpu().numpy().astype(np.float64)
    pot_new = -f1[:, K.SF_WTD] / 0.0165
    f1[:, K.SF_POTENTIAL] = pot_new - (g["pot_new"] - g["pot_old"])
    _set(env, f1, i1.cpu().numpy())
    obs, rew, done, info = env.step(torch.zeros(M, 8, device="cuda"))
    np.testing.assert_allclose(rew.cpu().numpy(), g["mj_rew"], rtol=0, atol=1e-4)   # north-star reward tolerance
    assert np.array_equal(done.cpu().numpy(), g["mj_done"])
