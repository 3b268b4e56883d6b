"""CPU-side checks of the C-ABI library: it builds, loads, and exports every symbol that
include/hrl_b200.h declares; config structs agree between C and ctypes; no compute calls."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as G
from hrl_pybullet_envs_b200 import _cabi, config as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    G.build()
    return _cabi.lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "hrl_b200.h")).read()
    declared = set(re.findall(r"\b(hrl_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_default_configs_match_oracle(lib):
    from oracle import oracle as O
    for kind in range(6):
        a = _cabi.default_config(kind, 7)
        b = O.default_config(kind, 7)
        assert bytes(a) == bytes(b), kind
        assert lib.hrl_obs_dim(C.byref(a)) == O.lib().hrlo_obs_dim(C.byref(b)) == K.obs_dim(a)
        assert lib.hrl_act_dim(C.byref(a)) == K.act_dim(a)


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _cabi.default_config(K.HRL_ANT_GATHER, 4)
    h = C.c_void_p()
    rc = lib.hrl_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.hrl_last_error()
    from hrl_pybullet_envs_b200 import VecEnv
    with pytest.raises(_cabi.HrlError):
        VecEnv("AntGatherBulletEnv-v0", 4)


def test_invalid_configs_rejected(lib):
    cfg = _cabi.default_config(K.HRL_ANT_GATHER, 4)
    cfg.n_bins = 99
    h = C.c_void_p()
    assert lib.hrl_create(C.byref(cfg), 0, C.byref(h)) == -1
    bad = K.HrlConfig()
    assert lib.hrl_default_config(42, 1, C.byref(bad)) == -1


def test_kwargs_mapping():
    cfg = _cabi.default_config(K.HRL_ANT_GATHER, 1)
    K.apply_kwargs(cfg, K.HRL_ANT_GATHER, dict(n_food=3, n_poison=2, n_bins=6, dying_cost=-5, world_size=(11, 13)))
    assert (cfg.n_food, cfg.n_poison, cfg.n_bins, cfg.dying_cost) == (3, 2, 6, -5.0)
    assert tuple(cfg.world_size) == (11.0, 13.0)
    with pytest.raises(TypeError):
        K.apply_kwargs(cfg, K.HRL_ANT_GATHER, dict(nonsense=1))
    K.apply_kwargs(cfg, K.HRL_ANT_GATHER, dict(use_sensor=False))      # get_abs_pos: xy of the nearest items
    assert cfg.use_sensor == 0 and K.obs_dim(cfg) == 26 + 2 * 3 + 2 * 2
    assert cfg.item_contacts == 1 and cfg.item_friction == pytest.approx(0.75)   # cubes are colliders by default
    K.apply_kwargs(cfg, K.HRL_ANT_GATHER, dict(item_contacts=False, robot_coll_dist=0))   # contact-based pickup needs them
    assert cfg.item_contacts == 1
    p = _cabi.default_config(K.HRL_POINT_GATHER, 1)
    K.apply_kwargs(p, K.HRL_POINT_GATHER, dict(use_sensor=False, n_bins=4))   # gather_base.py:170-187
    assert p.use_sensor == 0 and K.obs_dim(p) == 8 + 2 * 4 + 2 * 4
    with pytest.raises(NotImplementedError):                             # still outside the built scope: fail loudly
        K.apply_kwargs(_cabi.default_config(K.HRL_POINT_GATHER, 1), K.HRL_POINT_GATHER, dict(robot_coll_dist=0))
    f = _cabi.default_config(K.HRL_ANT_FLAGRUN, 1)
    K.apply_kwargs(f, K.HRL_ANT_FLAGRUN, dict(use_sensor=True, sensor_bins=6, switch_flag_on_collision=False))
    assert f.flag_use_sensor == 1 and f.flag_switch_on_collision == 0 and K.obs_dim(f) == 34
    with pytest.raises(AssertionError):                                  # ant_flagrun_env.py:17-18
        K.apply_kwargs(_cabi.default_config(K.HRL_ANT_FLAGRUN, 1), K.HRL_ANT_FLAGRUN, dict(max_target_dist=3.0))
    g = _cabi.default_config(K.HRL_ANT_FLAGRUN, 1)
    K.apply_kwargs(g, K.HRL_ANT_FLAGRUN, dict(manual_goal_creation=True))   # ant_flagrun_env.py:150-153
    assert g.flag_manual_goals == 1
    o = _cabi.default_config(K.HRL_ANT_FLAGRUN, 1)                       # ant_flagrun_env.py:59-69: the open stadium scene
    K.apply_kwargs(o, K.HRL_ANT_FLAGRUN, dict(enclosed=False))
    assert (o.has_walls, o.ground_z, o.n_scene_parts) == (0, 0.0, 3)
    o2 = _cabi.default_config(K.HRL_ANT_FLAGRUN, 1)                      # use_sensor keeps the walled arena (:60)
    K.apply_kwargs(o2, K.HRL_ANT_FLAGRUN, dict(enclosed=False, use_sensor=True))
    assert o2.has_walls == 1
    m = _cabi.default_config(K.HRL_ANT_MAZE, 1)
    K.apply_kwargs(m, K.HRL_ANT_MAZE, dict(targets=([1, 2], [3, 4]), tol=2.0, target_encoding=1))
    assert m.n_targets == 2 and m.targets[1][0] == 3.0 and m.tol == 2.0 and m.target_encoding == 1
    assert K.obs_dim(m) == 38
    K.apply_kwargs(m, K.HRL_ANT_MAZE, dict(sense_target=True, max_steps=50, targ_dist_rew=True))
    assert (m.sense_target, m.maze_max_steps, m.targ_dist_rew) == (1, 50, 1) and K.obs_dim(m) == 46
