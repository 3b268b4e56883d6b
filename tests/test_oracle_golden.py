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


@pytest.mark.parametrize("tag,kind,nbase,can_die", [("ant", K.HRL_ANT_GATHER, 26, 1), ("point", K.HRL_POINT_GATHER, 8, 0)])
def test_gather_step_task_layer(golden, tag, kind, nbase, can_die):
    """Whole reference AntGather/PointGather `step` task layer (pickups, respawn rule with
    replayed uniforms, sensor, alive/done, reward, info)."""
    g = golden("gather_step.npz")
    L = O.lib()
    cfg = O.default_config(kind, 1)
    nb = cfg.n_bins
    bad = 0
    for m in range(len(g[f"{tag}_state"])):
        st = g[f"{tag}_state"][m].copy()
        xyz = g[f"{tag}_xyz"][m]
        st[0] = np.float32(xyz[2] - (0.75 if tag == "ant" else 1.0))
        st = st.astype(np.float32).astype(np.float64)
        base = np.ascontiguousarray(np.concatenate([st[0:1], st[3:]]) if tag == "ant" else st)
        z_alive = st[0] + (0.75 if tag == "ant" else 1.0)
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
