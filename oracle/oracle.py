"""ctypes binding of the CPU oracle (oracle/hrl_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

from hrl_pybullet_envs_b200.config import (HrlConfig, HRL_STATE_F, HRL_STATE_I, HRL_MAX_ITEMS, ENV_IDS, apply_kwargs)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    out = os.path.join(_HERE, "_build", "libhrl_oracle.so")
    src = os.path.join(_HERE, "hrl_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return out


def lib(f32=False, count=False):
    """count=True: the C++ build whose `real` counts its arithmetic (oracle/flop_counter.hpp, `make count`)."""
    key = "count" if count else bool(f32)
    if key not in _LIBS:
        build()
        if count:
            subprocess.check_call(["make", "-C", _HERE, "-s", "count"], stdout=subprocess.DEVNULL)
        path = os.path.join(_HERE, "_build", "libhrl_oracle_count.so" if count else ("libhrl_oracle_f32.so" if f32 else "libhrl_oracle.so"))
        L = C.CDLL(path)
        L.hrlo_default_config.argtypes = [C.c_int32, C.c_int32, C.POINTER(HrlConfig)]
        L.hrlo_create.argtypes = [C.POINTER(HrlConfig), C.POINTER(C.c_void_p)]
        L.hrlo_destroy.argtypes = [C.c_void_p]
        L.hrlo_obs_dim.argtypes = [C.POINTER(HrlConfig)]
        L.hrlo_act_dim.argtypes = [C.POINTER(HrlConfig)]
        L.hrlo_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hrlo_step.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.hrlo_step_range.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        L.hrlo_observe.argtypes = [C.c_void_p, C.c_void_p]
        L.hrlo_substeps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.hrlo_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hrlo_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hrlo_gather_sensor_one.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hrlo_sense_walls_one.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_double, C.c_double,
                                           C.c_double, C.c_void_p]
        L.hrlo_quadrant.argtypes = [C.c_double, C.c_double]
        L.hrlo_find_intersection.argtypes = [C.c_void_p, C.c_void_p]
        L.hrlo_segment_intersection.argtypes = [C.c_void_p]
        L.hrlo_random_on_plane_replay.argtypes = [C.c_double] * 5 + [C.c_void_p, C.c_void_p]
        L.hrlo_gather_task_replay.argtypes = [C.POINTER(HrlConfig), C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.hrlo_scene_bounds.argtypes = [C.POINTER(HrlConfig), C.c_void_p]
        L.hrlo_mass_matrix.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hrlo_free_accel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hrlo_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.hrlo_rng_u4.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.hrlo_flag_goal.argtypes = [C.POINTER(HrlConfig), C.c_int, C.c_int, C.c_void_p]
        d = C.c_double
        L.hrlo_maze_target_sensor.argtypes = [C.c_int, d, d, C.c_int, C.c_void_p, d, d, d, d, d, d, C.c_void_p]
        L.hrlo_maze_target_sensor.restype = None
        L.hrlo_maze_task_replay.argtypes = [C.POINTER(HrlConfig), C.c_void_p, C.c_void_p, d, d, C.c_int, d, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_void_p]
        L.hrlo_maze_mj_task_replay.argtypes = [C.POINTER(HrlConfig), C.c_void_p, d, d, C.c_int, d, C.c_int, C.c_void_p, C.c_void_p]
        L.hrlo_flagrun_replay.argtypes = [C.POINTER(HrlConfig), C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.hrlo_point_state.argtypes = [C.c_void_p] * 4
        L.hrlo_point_state.restype = None
        L.hrlo_point_force.argtypes = [C.POINTER(HrlConfig), C.c_void_p, C.c_void_p]
        L.hrlo_point_force.restype = None
        L.hrlo_mj_reward.argtypes = [C.POINTER(HrlConfig), d, d, d, C.c_int, C.c_void_p]
        L.hrlo_mj_reward.restype = None
        _LIBS[key] = L
    return _LIBS[key]


def default_config(kind, num_envs, **kw):
    cfg = HrlConfig()
    rc = lib().hrlo_default_config(int(kind), int(num_envs), C.byref(cfg))
    if rc:
        raise ValueError("unknown env kind %r" % kind)
    seed = kw.pop("seed_", None)
    apply_kwargs(cfg, kind, kw)
    if seed is not None:
        cfg.seed = seed
    return cfg


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleVecEnv:
    """N envs stepped on the host by the C oracle; same call shapes as the CUDA VecEnv."""

    def __init__(self, cfg, f32=False, threads=1, count=False):
        self.L = lib(f32, count)
        self.cfg = cfg.copy()
        self.real = np.float32 if f32 else np.float64
        self.h = C.c_void_p()
        rc = self.L.hrlo_create(C.byref(self.cfg), C.byref(self.h))
        if rc:
            raise ValueError("hrlo_create failed: %d" % rc)
        self.N = cfg.num_envs
        self.D = self.L.hrlo_obs_dim(C.byref(self.cfg))
        self.A = self.L.hrlo_act_dim(C.byref(self.cfg))
        self.threads = max(1, int(threads))

    @classmethod
    def make(cls, env_id, num_envs, seed=None, f32=False, threads=1, env_seed=None, **kw):
        """Same seeding rule as the CUDA VecEnv: `seed` keys the per-env streams and, for Flagrun, an explicit one is
        also the reference's ctor kwarg that seeds the shared goal stream (ant_flagrun_env.py:16,39)."""
        from hrl_pybullet_envs_b200.config import HRL_ANT_FLAGRUN
        cfg = default_config(ENV_IDS[env_id], num_envs, **kw)
        cfg.seed = int(seed or 0)
        if ENV_IDS[env_id] == HRL_ANT_FLAGRUN and seed is not None:
            cfg.flag_seed = int(seed)
        if env_seed is not None:
            cfg.seed = int(env_seed)
        return cls(cfg, f32=f32, threads=threads)

    def close(self):
        if self.h:
            self.L.hrlo_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, mask=None):
        obs = np.zeros((self.N, self.D), dtype=self.real)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.hrlo_reset(self.h, _p(m), _p(obs))
        return obs

    def step(self, actions, want_terminal=False):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.N, self.A)
        obs = np.zeros((self.N, self.D), dtype=self.real)
        rew = np.zeros(self.N, dtype=self.real)
        done = np.zeros(self.N, dtype=np.uint8)
        info = np.zeros((self.N, 4), dtype=self.real)
        term = np.zeros((self.N, self.D), dtype=self.real) if want_terminal else None
        if self.threads == 1:
            self.L.hrlo_step(self.h, _p(a), _p(obs), _p(rew), _p(done), _p(info), _p(term))
        else:  # envs are independent: shard across host threads (ctypes releases the GIL)
            edges = np.linspace(0, self.N, self.threads + 1).astype(int)
            ts = [threading.Thread(target=self.L.hrlo_step_range,
                                   args=(self.h, int(edges[i]), int(edges[i + 1]), _p(a), _p(obs), _p(rew), _p(done),
                                         _p(info), _p(term))) for i in range(self.threads)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        out = (obs, rew, done.astype(bool), info)
        return out + (term,) if want_terminal else out

    def observe(self):
        obs = np.zeros((self.N, self.D), dtype=self.real)
        self.L.hrlo_observe(self.h, _p(obs))
        return obs

    def substeps(self, actions, n_sub):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.N, self.A)
        self.L.hrlo_substeps(self.h, _p(a), int(n_sub))

    def get_state(self):
        f = np.zeros((self.N, HRL_STATE_F), dtype=self.real)
        i = np.zeros((self.N, HRL_STATE_I), dtype=np.int32)
        self.L.hrlo_get_state(self.h, _p(f), _p(i))
        return f, i

    def set_state(self, f, i):
        f = np.ascontiguousarray(f, dtype=self.real).reshape(self.N, HRL_STATE_F)
        i = np.ascontiguousarray(i, dtype=np.int32).reshape(self.N, HRL_STATE_I)
        self.L.hrlo_set_state(self.h, _p(f), _p(i))

    def inverse_mass_matrix(self, e=0):
        M = np.zeros((14, 14), dtype=self.real)
        rc = self.L.hrlo_mass_matrix(self.h, int(e), _p(M))
        if rc:
            raise RuntimeError("ABA failed")
        return M

    def free_accel(self, e=0, tau=None):
        tau = np.zeros(8, dtype=self.real) if tau is None else np.ascontiguousarray(tau, dtype=self.real)
        ud = np.zeros(14, dtype=self.real)
        rc = self.L.hrlo_free_accel(self.h, int(e), _p(tau), _p(ud))
        if rc:
            raise RuntimeError("ABA failed")
        return ud

    def stats(self):
        s = np.zeros(3, dtype=np.float64)
        self.L.hrlo_stats(self.h, _p(s))
        return {"contacts_per_substep": s[0] / max(s[2], 1), "limit_rows_per_substep": s[1] / max(s[2], 1),
                "substeps": s[2]}


# ---- stateless helpers used by the golden-vector tests ------------------------------------
def gather_sensor(n_bins, sensor_range, span, xy, yaw, items, n_food=8, n_poison=8):
    L = lib()
    items = np.ascontiguousarray(items, dtype=np.float64).reshape(-1, HRL_MAX_ITEMS, 2)
    M = items.shape[0]
    xy = np.asarray(xy, dtype=np.float64).reshape(M, 2)
    yaw = np.asarray(yaw, dtype=np.float64).reshape(M)
    food = np.zeros((M, n_bins)); poison = np.zeros((M, n_bins)); bins = np.zeros((M, HRL_MAX_ITEMS), dtype=np.int32)
    for m in range(M):
        it = np.ascontiguousarray(items[m])
        L.hrlo_gather_sensor_one(n_bins, sensor_range, span, xy[m, 0], xy[m, 1], yaw[m], _p(it), n_food, n_poison, None,
                                 _p(food[m]), _p(poison[m]), _p(bins[m]))
    return food, poison, bins


def sense_walls(n_bins, span, rng, bounds, xy, yaw):
    L = lib()
    bounds = np.ascontiguousarray(bounds, dtype=np.float64)
    xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    yaw = np.asarray(yaw, dtype=np.float64).reshape(-1)
    out = np.zeros((xy.shape[0], n_bins))
    for m in range(xy.shape[0]):
        L.hrlo_sense_walls_one(n_bins, span, rng, bounds.shape[0], _p(bounds), xy[m, 0], xy[m, 1], yaw[m], _p(out[m]))
    return out


def scene_bounds(cfg):
    b = np.zeros((7, 4))
    n = lib().hrlo_scene_bounds(C.byref(cfg), _p(b))
    return b[:n]
