"""The CPU oracle against golden vectors produced by the REFERENCE's own Python task logic
(tests/golden/make_golden.py).  This is what pins the oracle (task layer)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from hrl_pybullet_envs_b200 import config as K

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_intersection_utils(golden):
    g = golden("intersection.npz")
    L = O.lib()
    for p, want in zip(g["seg_in"], g["seg_out"]):
        p = np.ascontiguousarray(p)
        assert bool(L.hrlo_segment_intersection(O._p(p))) == bool(want)
    for p, has, xy in zip(g["seg_in"], g["inf_has"], g["inf_xy"]):
        p = np.ascontiguousarray(p); out = np.zeros(2)
        assert bool(L.hrlo_find_intersection(O._p(p), O._p(out))) == bool(has)
        if has:
            assert np.array_equal(out, xy)  # same IEEE expression => bit-identical
    for p, q in zip(g["quad_in"], g["quad_out"]):
        assert L.hrlo_quadrant(p[0], p[1]) == q


@pytest.mark.parametrize("tag,n_bins", [("ant", 10), ("point", 5)])
def test_gather_sensor(golden, tag, n_bins):
    g = golden("gather_sensor.npz")
    food, poison, bins = O.gather_sensor(n_bins, 20.0, np.pi, g["xy"], g["yaw"], g["objs"])
    # bins bit-exact <=> zero pattern identical; intensities: same double expression => exact
    assert np.array_equal(food != 0, g[f"food_{tag}"] != 0)
    assert np.array_equal(poison != 0, g[f"poison_{tag}"] != 0)
    np.testing.assert_allclose(food, g[f"food_{tag}"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(poison, g[f"poison_{tag}"], rtol=0, atol=1e-15)
    # SURVEY.md 8c(1) known answer
    if tag == "ant":
        assert abs(food[0, 6] - 0.951) < 1e-6 and abs(poison[0, 8] - 0.756) < 1e-6  # inputs are f32-rounded


@pytest.mark.parametrize("tag", ["maze", "flagrun"])
def test_sense_walls(golden, tag):
    g = golden("sense_walls.npz")
    kind = K.HRL_ANT_MAZE if tag == "maze" else K.HRL_ANT_FLAGRUN
    cfg = O.default_config(kind, 1)
    b = O.scene_bounds(cfg)
    assert np.array_equal(b, g[f"{tag}_bounds"])  # bound lines, in the reference's order (8c(3))
    full = O.sense_walls(10, 2 * np.pi, 5.0, b, g[f"{tag}_xy"], g[f"{tag}_yaw"])
    half = O.sense_walls(8, np.pi, 4.0, b, g[f"{tag}_xy"], g[f"{tag}_yaw"])
    np.testing.assert_allclose(full, g[f"{tag}_full10_r5"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(half, g[f"{tag}_pi8_r4"], rtol=0, atol=1e-12)
    if tag == "maze":
        want = [0.258359213500, 0.369122665457, 0.369122665457, 0.011145618000, 0.2, 0.011145618000,
                0.369122665457, 0.369122665457, 0.258359213500, 0.4]
        np.testing.assert_allclose(full[0], want, atol=1e-9)


def test_random_on_plane(golden):
    g = golden("random_on_plane.npz")
    L = O.lib()
    for u, avoid, pos, used in zip(g["u"], g["avoid"], g["pos"], g["used"]):
        out = np.zeros(2); u = np.ascontiguousarray(u)
        n = L.hrlo_random_on_plane_replay(15.0, 15.0, 2.0, avoid[0], avoid[1], O._p(u), O._p(out))
        assert n == used
        np.testing.assert_allclose(out, pos[:2], rtol=0, atol=1e-12)
        assert pos[2] == 0.1
    # episode_restart: 16 items, re-randomised avoiding (0,0) (second restart draws 16 accepted positions)
    u2 = g["u2"]; i = int(g["used1"]); got = []
    for k in range(16):
        out = np.zeros(2); uu = np.ascontiguousarray(u2[i:])
        i += L.hrlo_random_on_plane_replay(15.0, 15.0, 2.0, 0.0, 0.0, O._p(uu), O._p(out))
        got.append(out.copy())
    assert i == int(g["used2"])
    np.testing.assert_allclose(np.array(got), g["restart2"][:, :2], atol=1e-12)
    assert list(g["rew"]) == [1, -1, 0] and int(g["rew_norespawn"]) == 1
    assert list(g["parked"]) == [100, 0, -10]


@pytest.mark.parametrize("tag,kind,nbase,can_die", [("ant", K.HRL_ANT_GATHER, 26, 1), ("point", K.HRL_POINT_GATHER, 8, 0),
                                                    ("antabs", K.HRL_ANT_GATHER, 26, 1), ("pointabs", K.HRL_POINT_GATHER, 8, 0)])
def test_gather_step_task_layer(golden, tag, kind, nbase, can_die):
    """Whole reference AntGather/PointGather `step` task layer (pickups, respawn rule with
    replayed uniforms, sensor, alive/done, reward, info)."""
    g = golden("gather_step_pointabs.npz" if tag == "pointabs" else "gather_step.npz")
    L = O.lib()
    cfg = O.default_config(kind, 1)
    if tag.endswith("abs"):  # use_sensor=False: xy of the n_bins nearest food / poison items (ant_gather_env.py:179-196)
        cfg.use_sensor = 0; cfg.n_bins = 5 if tag == "antabs" else 4
    nb = K.food_obs_dim(cfg) // 2
    bad = 0
    for m in range(len(g[f"{tag}_state"])):
        st = g[f"{tag}_state"][m].copy()
        xyz = g[f"{tag}_xyz"][m]
        st[0] = np.float32(xyz[2] - (0.75 if tag.startswith("ant") else 1.0))
        st = st.astype(np.float32).astype(np.float64)
        base = np.ascontiguousarray(np.concatenate([st[0:1], st[3:]]) if tag.startswith("ant") else st)
        z_alive = st[0] + (0.75 if tag.startswith("ant") else 1.0)
        items = np.ascontiguousarray(g[f"{tag}_objs"][m].reshape(-1).copy())
        u = np.ascontiguousarray(g[f"{tag}_u"][m])
        obs = np.zeros(nbase + 2 * nb); rdi = np.zeros(4); used = O.C.c_int(0)
        xyz_in = np.array([xyz[0], xyz[1], z_alive])
        L.hrlo_gather_task_replay(O.C.byref(cfg), O._p(base), nbase, O._p(xyz_in), float(g[f"{tag}_rpy"][m, 2]), can_die,
                                  O._p(items), O._p(u), O._p(obs), O._p(rdi), O.C.byref(used))
        want = g[f"{tag}_obs"][m]
        assert used.value == g[f"{tag}_used"][m]
        np.testing.assert_allclose(items.reshape(16, 2), g[f"{tag}_new_objs"][m], atol=1e-12)
        both_nan = np.isnan(obs) & np.isnan(want)
        np.testing.assert_allclose(np.where(both_nan, 0, obs), np.where(both_nan, 0, want), rtol=0, atol=1e-12)
        assert rdi[0] == g[f"{tag}_rew"][m] and bool(rdi[1]) == bool(g[f"{tag}_done"][m])
        assert rdi[2] == g[f"{tag}_food_rew"][m] and rdi[3] == g[f"{tag}_dead_rew"][m]
    assert bad == 0


def test_registry():
    with open(os.path.join(GOLDEN, "registry.json")) as f:
        reg = json.load(f)
    ids = {r["id"] for r in reg}
    assert ids == {"AntGatherBulletEnv-v0", "AntMazeMjEnv-v0", "AntMazeBulletEnv-v0", "AntFlagrunBulletEnv-v0",
                   "PointGatherBulletEnv-v0"}
    assert all(r["max_episode_steps"] == 2000 for r in reg)
    assert ids <= set(K.ENV_IDS)


# ---------------------------------------------------------------- maze goal observations (maze_target.npz)
def test_maze_target_vec_and_sensor(golden):
    """get_target_vec_obs (both encodings) and get_target_sensor_obs (range, box occlusion, bin)
    of ant_maze_bullet_env.py:123-178 on 300 poses."""
    g = golden("maze_target.npz")
    L = O.lib()
    cfg = O.default_config(K.HRL_ANT_MAZE, 1)
    box = np.ascontiguousarray(O.scene_bounds(cfg)[4:7])
    assert box.shape == (3, 4)
    for m in range(len(g["xy"])):
        xy = g["xy"][m].astype(np.float64); yaw = float(g["yaw"][m]); tgt = g["targets"][g["tid"][m]].astype(np.float64)
        rd = np.zeros(10)
        L.hrlo_maze_target_sensor(10, 2 * np.pi, 5.0, 3, O._p(box), xy[0], xy[1], yaw, tgt[0], tgt[1], float(g["wtd"][m]), O._p(rd))
        np.testing.assert_allclose(rd, g["sensor"][m], rtol=0, atol=1e-12)
        for enc, key in ((0, "vec_normed"), (1, "vec_angle")):
            c = cfg.copy(); c.target_encoding = enc
            obs = np.zeros(38); rd2 = np.zeros(2)
            L.hrlo_maze_task_replay(O.C.byref(c), O._p(np.zeros(28)), O._p(np.ascontiguousarray(xy)), yaw, 0.0, 0, float(g["wtd"][m]),
                                    O._p(np.ascontiguousarray(tgt)), 0, O._p(obs), O._p(rd2))
            np.testing.assert_allclose(obs[26:28], g[key][m], rtol=0, atol=1e-12)
    assert g["sensor"][0][7] == pytest.approx(0.6) and not g["sensor"][1].any()  # SURVEY.md 8c known answers


# ---------------------------------------------------------------- whole Maze step task layer (maze_step*.npz)
def _maze_cfg(variant):
    cfg = O.default_config(K.HRL_ANT_MAZE, 1)
    for k, v in variant.items():
        setattr(cfg, k, v)
    return cfg


@pytest.mark.parametrize("tag", ["", "_sense_target", "_max_steps", "_targ_dist", "_angle_nowalls"])
def test_maze_step_task_layer(golden, tag):
    """AntMazeBulletEnv.step with a stub inner walker step: observation (goal part + lidar), reward, done
    (ant_maze_bullet_env.py:63-97) for the default kwargs and the non-default ones."""
    g = golden("maze_step.npz")
    if "obs" + tag not in g:
        pytest.skip("variant not in the fixture")
    L = O.lib()
    variant = {"": {}, "_sense_target": {"sense_target": 1},
               "_max_steps": {"maze_max_steps": 7, "done_at_target": 0},
               "_targ_dist": {"targ_dist_rew": 1, "inner_rew_weight": 0.5},
               "_angle_nowalls": {"target_encoding": 1, "sense_walls": 0}}[tag]
    cfg = _maze_cfg(variant)
    D = L.hrlo_obs_dim(O.C.byref(cfg))
    tb = g["t_before" + tag] if "t_before" + tag in g else np.zeros(len(g["xy"]), int)
    for m in range(len(g["xy"])):
        obs = np.zeros(D); rd = np.zeros(2)
        tgt = np.ascontiguousarray(g["targets"][g["tid"][m]].astype(np.float64))
        L.hrlo_maze_task_replay(O.C.byref(cfg), O._p(np.ascontiguousarray(g["ant_obs"][m].astype(np.float64))),
                                O._p(np.ascontiguousarray(g["xy"][m].astype(np.float64))), float(g["yaw"][m]),
                                float(g["inner_rew"][m]), int(g["inner_done"][m]), float(g["wtd"][m]), O._p(tgt), int(tb[m]),
                                O._p(obs), O._p(rd))
        want = g["obs" + tag][m]
        assert want.shape == (D,)
        np.testing.assert_allclose(obs, want, rtol=0, atol=1e-12)
        assert rd[0] == pytest.approx(g["rew" + tag][m], abs=1e-12) and bool(rd[1]) == bool(g["done" + tag][m])


# ---------------------------------------------------------------- whole MazeMj step task layer (maze_mj_step.npz)
@pytest.mark.parametrize("tag", ["", "_weight"])
def test_maze_mj_step_task_layer(golden, tag):
    """AntMazeMjEnv.step / _get_obs with a stub inner AntMjEnv step (ant_maze_mj_env.py:57-79): the 60-d observation
    [mj 29 | walls 10 | 0 x 10 | 0 x 10 | t * 0.001 with t taken BEFORE the increment], reward and done."""
    g = golden("maze_mj_step.npz")
    L = O.lib()
    cfg = O.default_config(K.HRL_ANT_MAZE_MJ, 1)
    for k, v in {"": {}, "_weight": {"inner_rew_weight": 0.5, "tol": 2.5}}[tag].items():
        setattr(cfg, k, v)
    assert L.hrlo_obs_dim(O.C.byref(cfg)) == 60
    assert np.array_equal(g["targets"], np.array([list(cfg.targets[i]) for i in range(cfg.n_targets)]))   # ant_maze_mj_env.py:13-14
    for m in range(len(g["mj_obs"])):
        obs = np.zeros(60); rd = np.zeros(2)
        L.hrlo_maze_mj_task_replay(O.C.byref(cfg), O._p(np.ascontiguousarray(g["mj_obs"][m].astype(np.float64))), float(g["yaw"][m]),
                                   float(g["inner_rew"][m]), int(g["inner_done"][m]), float(g["wtd"][m]), int(g["t_before"][m]),
                                   O._p(obs), O._p(rd))
        np.testing.assert_allclose(obs, g["obs" + tag][m], rtol=0, atol=1e-12)
        assert rd[0] == pytest.approx(g["rew" + tag][m], abs=1e-12) and bool(rd[1]) == bool(g["done" + tag][m])
        assert g["t_after" + tag][m] == g["t_before"][m] + 1
    assert g["done"].any() and not g["done"].all() and (g["rew"] == 1).any()


# ---------------------------------------------------------------- Flagrun step sequences (flagrun_step.npz)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_flagrun_step_sequence(golden, tag):
    """AntFlagrunBulletEnv.step bookkeeping (ant_flagrun_env.py:162-204) on a scripted sequence: +5000 once
    per goal, goals popped from the END of the list, 200-step timeout, IndexError -> done, the
    switch_flag_on_collision kwarg (sequence c)."""
    g = golden("flagrun_step.npz")
    if tag + "_rew" not in g:
        pytest.skip("variant not in the fixture")
    L = O.lib()
    cfg = O.default_config(K.HRL_ANT_FLAGRUN, 1)
    cfg.flag_max_targets = int(g[tag + "_n_goals"]); cfg.flag_timeout = int(g[tag + "_timeout"])
    if tag + "_switch" in g:
        cfg.flag_switch_on_collision = int(g[tag + "_switch"])
    goals = np.ascontiguousarray(g[tag + "_goals0"].astype(np.float64))
    T = len(g["wtd"])
    rew = np.zeros(T); done = np.zeros(T, np.int32); tg = np.zeros((T, 2)); since = np.zeros(T, np.int32); rw = np.zeros(T, np.int32)
    n = L.hrlo_flagrun_replay(O.C.byref(cfg), O._p(goals), T, O._p(np.ascontiguousarray(g["wtd"].astype(np.float64))),
                              O._p(np.ascontiguousarray(g["inner_r"].astype(np.float64))), O._p(rew), O._p(done), O._p(tg),
                              O._p(since), O._p(rw))
    assert n == len(g[tag + "_rew"])
    np.testing.assert_allclose(rew[:n], g[tag + "_rew"], rtol=0, atol=1e-9)
    assert np.array_equal(done[:n].astype(bool), g[tag + "_done"])
    np.testing.assert_allclose(tg[:n], g[tag + "_target"], rtol=0, atol=1e-12)
    assert np.array_equal(since[:n], g[tag + "_since"]) and np.array_equal(rw[:n].astype(bool), g[tag + "_rewarded"])


def test_flagrun_goal_rule(golden):
    """create_target (ant_flagrun_env.py:71-78): the reference's MT19937 stream is not reproduced (DESIGN.md
    section 2), the placement rule is: uniform on the size^2 square, never closer than 0.5 to the origin."""
    g = golden("flagrun_goals.npz")
    for key in ("goals1", "goals2"):
        G = g[key]
        assert G.shape == (100, 2) and (np.abs(G) <= 5).all() and (np.linalg.norm(G, axis=1) >= 0.5).all()
    cfg = O.default_config(K.HRL_ANT_FLAGRUN, 1)
    ours = np.zeros((100, 2))
    for j in range(100):
        gj = np.zeros(2); O.lib().hrlo_flag_goal(O.C.byref(cfg), 3, j, O._p(gj)); ours[j] = gj
    assert (np.abs(ours) <= 5).all() and (np.linalg.norm(ours, axis=1) >= 0.5).all()
    assert abs(ours.mean()) < 1.0 and 2.0 < ours.std() < 3.6  # U(-5,5): std 2.89


# ---------------------------------------------------------------- PointBot + AntMjEnv reward (robots.npz)
def test_pointbot_and_mj_reward(golden):
    g = golden("robots.npz")
    L = O.lib()
    cfgp = O.default_config(K.HRL_POINT_GATHER, 1); cfgm = O.default_config(K.HRL_ANT_MJ, 1)
    for m in range(len(g["xyz"])):
        out = np.zeros(8)
        L.hrlo_point_state(O._p(np.ascontiguousarray(g["xyz"][m].astype(np.float64))), O._p(np.ascontiguousarray(g["rpy"][m].astype(np.float64))),
                           O._p(np.ascontiguousarray(g["vel"][m].astype(np.float64))), O._p(out))
        np.testing.assert_allclose(out, g["point_state"][m], rtol=0, atol=2e-6)  # the reference returns float32
        f = np.zeros(3)
        L.hrlo_point_force(O.C.byref(cfgp), O._p(np.ascontiguousarray(g["act"][m].astype(np.float64))), O._p(f))
        np.testing.assert_allclose(f, g["force"][m], rtol=1e-6, atol=1e-4)
        rd = np.zeros(2)
        L.hrlo_mj_reward(O.C.byref(cfgm), float(g["mj_state"][m, 2]), float(g["pot_old"][m]), float(g["pot_new"][m]), int(g["jal"][m]), O._p(rd))
        assert rd[0] == pytest.approx(g["mj_rew"][m], abs=1e-5) and bool(rd[1]) == bool(g["mj_done"][m])
