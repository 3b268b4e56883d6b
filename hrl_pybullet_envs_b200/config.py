"""ctypes mirror of ``hrl_config`` (include/hrl_b200.h) and the mapping from the reference's
constructor kwargs to it.

Reference kwargs (paths relative to /root/reference/hrl_pybullet_envs):
  AntGatherBulletEnv   envs/gather/ant_gather_env.py:16-29
  PointGatherBulletEnv envs/gather/point_gather_env.py:8-21
  AntMazeBulletEnv     envs/ant_maze/ant_maze_bullet_env.py:23-25
  AntMazeMjEnv         envs/ant_maze/ant_maze_mj_env.py:23-25
  AntFlagrunBulletEnv  envs/ant_flagrun/ant_flagrun_env.py:14-16
  AntMjEnv             envs/MjAnt.py:31-34 (no kwargs)
"""
import ctypes as C
import math

HRL_ANT_GATHER, HRL_ANT_MAZE, HRL_ANT_FLAGRUN, HRL_ANT_MJ, HRL_POINT_GATHER, HRL_ANT_MAZE_MJ = range(6)
HRL_MAX_ITEMS = 16
HRL_MAX_TARGETS = 8
HRL_MAX_BINS = 16
HRL_STATE_F = 72
HRL_STATE_I = 8

# float-state offsets (enum HRL_SF_* in include/hrl_b200.h)
SF_POS, SF_QUAT, SF_LINVEL, SF_ANGVEL, SF_Q, SF_QD = 0, 3, 7, 10, 13, 21
SF_INITIAL_Z, SF_POTENTIAL, SF_TARGET, SF_WTD, SF_FEET, SF_ITEMS = 29, 30, 31, 33, 34, 38
SF_RETURN, SF_RETURN_SUM = 70, 71
SI_T, SI_EPISODE, SI_STEPS, SI_GOALS_LEFT, SI_SINCE, SI_REWARDED, SI_GOAL_GEN = range(7)


class HrlConfig(C.Structure):
    _fields_ = [
        ("env_kind", C.c_int32), ("num_envs", C.c_int32), ("seed", C.c_uint64),
        ("env_index_offset", C.c_int32), ("max_episode_steps", C.c_int32), ("auto_reset", C.c_int32),
        ("gravity", C.c_float), ("dt", C.c_float), ("substeps", C.c_int32), ("solver_iters", C.c_int32),
        ("contact_erp", C.c_float), ("limit_erp", C.c_float), ("lin_damping", C.c_float), ("ang_damping", C.c_float),
        ("friction", C.c_float), ("limit_max_impulse", C.c_float), ("max_coord_vel", C.c_float),
        ("contact_margin", C.c_float), ("torque_scale", C.c_float), ("torque_first_substep_only", C.c_int32),
        ("world_size", C.c_float * 2), ("ground_z", C.c_float), ("has_walls", C.c_int32), ("has_box", C.c_int32),
        ("box_lo", C.c_float * 3), ("box_hi", C.c_float * 3), ("start_pos", C.c_float * 3),
        ("n_scene_parts", C.c_int32), ("scene_parts_sum", C.c_float * 2),
        ("n_food", C.c_int32), ("n_poison", C.c_int32), ("n_bins", C.c_int32),
        ("sensor_range", C.c_float), ("sensor_span", C.c_float), ("robot_coll_dist", C.c_float),
        ("robot_object_spacing", C.c_float), ("dying_cost", C.c_float), ("respawn", C.c_int32), ("use_sensor", C.c_int32),
        ("n_targets", C.c_int32), ("targets", (C.c_float * 2) * HRL_MAX_TARGETS), ("tol", C.c_float),
        ("done_at_target", C.c_int32), ("inner_rew_weight", C.c_float), ("target_encoding", C.c_int32),
        ("sense_walls", C.c_int32), ("flag_max_targets", C.c_int32), ("flag_timeout", C.c_int32),
        ("flag_size", C.c_float), ("goal_reach_rew", C.c_float), ("flag_seed", C.c_uint64),
        ("electricity_cost", C.c_float), ("stall_torque_cost", C.c_float), ("joints_at_limit_cost", C.c_float),
        ("sense_target", C.c_int32), ("maze_max_steps", C.c_int32), ("targ_dist_rew", C.c_int32),
        ("flag_use_sensor", C.c_int32), ("flag_switch_on_collision", C.c_int32), ("flag_max_target_dist", C.c_float),
        ("item_contacts", C.c_int32), ("item_friction", C.c_float), ("item_half", C.c_float), ("item_z", C.c_float),
        ("flag_manual_goals", C.c_int32),
    ]

    def copy(self):
        c = HrlConfig()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(HrlConfig))
        return c


ENV_IDS = {
    "AntGatherBulletEnv-v0": HRL_ANT_GATHER,
    "AntMazeBulletEnv-v0": HRL_ANT_MAZE,
    "AntFlagrunBulletEnv-v0": HRL_ANT_FLAGRUN,
    "AntMjBulletEnv-v0": HRL_ANT_MJ,       # README.md:13 / BASELINE.json name; class AntMjEnv
    "AntMjEnv-v0": HRL_ANT_MJ,
    "PointGatherBulletEnv-v0": HRL_POINT_GATHER,
    "AntMazeMjEnv-v0": HRL_ANT_MAZE_MJ,
}

_UNSUPPORTED = "kwarg %s=%r is outside the hot-path scope of this build (SURVEY.md 8f item 3)"


def apply_kwargs(cfg, kind, kw):
    """Apply the reference's ctor kwargs (same names, same defaults) to a default config.

    kwargs that only affect rendering / debug drawing are accepted and ignored, like the
    reference ignores them headless; kwargs selecting code paths that are not built raise.
    """
    kw = dict(kw)
    for k in ("render", "debug"):
        kw.pop(k, None)
    if kind in (HRL_ANT_GATHER, HRL_POINT_GATHER):
        if "world_size" in kw:
            ws = kw.pop("world_size")
            cfg.world_size[0], cfg.world_size[1] = float(ws[0]), float(ws[1])
        for name in ("n_food", "n_poison", "n_bins"):
            if name in kw:
                setattr(cfg, name, int(kw.pop(name)))
        for name in ("sensor_range", "sensor_span", "robot_coll_dist", "dying_cost"):
            if name in kw:
                setattr(cfg, name, float(kw.pop(name)))
        if "robot_object_spacing" in kw:
            cfg.robot_object_spacing = float(kw.pop("robot_object_spacing"))
        if "respawn" in kw:
            cfg.respawn = int(bool(kw.pop("respawn")))
        if "use_sensor" in kw:
            cfg.use_sensor = int(bool(kw.pop("use_sensor")))  # False: get_abs_pos, ant_gather_env.py:179-196 / gather_base.py:170-187
        if "item_contacts" in kw:      # extension kwarg: switch the cube colliders (default on for AntGather) off / on
            cfg.item_contacts = int(bool(kw.pop("item_contacts")))
            if kind == HRL_POINT_GATHER and cfg.item_contacts:
                raise NotImplementedError(_UNSUPPORTED % ("item_contacts", True))
        if cfg.robot_coll_dist <= 0:  # contact-based pickup (ant_gather_env.py:113-116): the cubes must be colliders
            if kind == HRL_POINT_GATHER:
                raise NotImplementedError(_UNSUPPORTED % ("robot_coll_dist", cfg.robot_coll_dist))
            cfg.item_contacts = 1
    elif kind in (HRL_ANT_MAZE, HRL_ANT_MAZE_MJ):
        if "n_bins" in kw:
            cfg.n_bins = int(kw.pop("n_bins"))
        if "sensor_range" in kw:
            cfg.sensor_range = float(kw.pop("sensor_range"))
        if "sensor_span" in kw:
            cfg.sensor_span = float(kw.pop("sensor_span"))
        if "targets" in kw:
            t = kw.pop("targets")
            if len(t) > HRL_MAX_TARGETS:
                raise ValueError("at most %d targets" % HRL_MAX_TARGETS)
            cfg.n_targets = len(t)
            for i, (x, y) in enumerate(t):
                cfg.targets[i][0], cfg.targets[i][1] = float(x), float(y)
        if "target_encoding" in kw:
            cfg.target_encoding = int(getattr(kw["target_encoding"], "value", kw.pop("target_encoding")))
            kw.pop("target_encoding", None)
        if "tol" in kw:
            cfg.tol = float(kw.pop("tol"))
        if "inner_rew_weight" in kw:
            cfg.inner_rew_weight = float(kw.pop("inner_rew_weight"))
        if kind == HRL_ANT_MAZE:
            if "sense_walls" in kw:
                cfg.sense_walls = int(bool(kw.pop("sense_walls")))
            if "done_at_target" in kw:
                cfg.done_at_target = int(bool(kw.pop("done_at_target")))
            if "sense_target" in kw:
                cfg.sense_target = int(bool(kw.pop("sense_target")))
            if "max_steps" in kw:
                cfg.maze_max_steps = int(kw.pop("max_steps"))
            if "targ_dist_rew" in kw:
                cfg.targ_dist_rew = int(bool(kw.pop("targ_dist_rew")))
    elif kind == HRL_ANT_FLAGRUN:
        if "size" in kw:
            cfg.flag_size = float(kw.pop("size"))
            cfg.world_size[0] = cfg.world_size[1] = cfg.flag_size + 2  # ant_flagrun_env.py:62
            cfg.scene_parts_sum[0] = -cfg.world_size[0] / 2
        if "tolerance" in kw:
            cfg.tol = float(kw.pop("tolerance"))
        if "max_targets" in kw:
            cfg.flag_max_targets = int(kw.pop("max_targets"))
        if "timeout" in kw:
            cfg.flag_timeout = int(kw.pop("timeout"))
        if "max_target_dist" in kw:
            cfg.flag_max_target_dist = float(kw.pop("max_target_dist"))
        # ant_flagrun_env.py:17-18: exactly one of max_targets / max_target_dist drives the goals
        if not ((cfg.flag_max_target_dist == 0 and cfg.flag_max_targets > 0) or
                (cfg.flag_max_targets <= 0 and cfg.flag_max_target_dist > 0)):
            raise AssertionError("cannot have both max_targets and max_target_dist set at the same time")
        if "use_sensor" in kw:
            cfg.flag_use_sensor = int(bool(kw.pop("use_sensor")))
        if "switch_flag_on_collision" in kw:
            cfg.flag_switch_on_collision = int(bool(kw.pop("switch_flag_on_collision")))
        if "sensor_bins" in kw:
            cfg.n_bins = int(kw.pop("sensor_bins"))
        if "sensor_span" in kw:
            cfg.sensor_span = float(kw.pop("sensor_span"))
        if "sensor_range" in kw:
            cfg.sensor_range = float(kw.pop("sensor_range"))
        if "manual_goal_creation" in kw:   # ant_flagrun_env.py:150-153: reset() creates no goals; VecEnv.set_target / create_targets
            cfg.flag_manual_goals = int(bool(kw.pop("manual_goal_creation")))
        enclosed = bool(kw.pop("enclosed", True))
        if not (enclosed or cfg.flag_use_sensor):
            # ant_flagrun_env.py:59-69: neither enclosed nor sensing -> the stock pybullet_envs stadium scene [3P-MEM]:
            # an open ground plane at z = 0 (stadium_no_collision.sdf), no walls.  Its three SDF bodies sit at the origin
            # and join the robot's parts like the arena bodies do (quirk Q1): 3 scene parts, xy sum (0, 0) [3P-MEM L].
            cfg.has_walls = 0
            cfg.ground_z = 0.0
            cfg.n_scene_parts = 3
            cfg.scene_parts_sum[0] = cfg.scene_parts_sum[1] = 0.0
        if cfg.flag_max_targets > 127:
            raise ValueError("max_targets must be <= 127")
    if kw:
        raise TypeError("unexpected keyword arguments: %s" % sorted(kw))
    if cfg.n_bins > HRL_MAX_BINS or cfg.n_bins < 1:
        raise ValueError("n_bins must be in [1, %d]" % HRL_MAX_BINS)
    if cfg.n_food > 8 or cfg.n_poison > 8 or cfg.n_food < 0 or cfg.n_poison < 0:
        raise ValueError("n_food / n_poison must be in [0, 8]")
    return cfg


def food_obs_dim(cfg):
    if cfg.use_sensor:
        return 2 * cfg.n_bins
    return 2 * min(cfg.n_bins, cfg.n_food) + 2 * min(cfg.n_bins, cfg.n_poison)


def obs_dim(cfg):
    k = cfg.env_kind
    return {HRL_ANT_GATHER: 26 + food_obs_dim(cfg),
            HRL_ANT_MAZE: 26 + (cfg.n_bins if cfg.sense_target else 2) + (cfg.n_bins if cfg.sense_walls else 0),
            HRL_ANT_FLAGRUN: 28 + (cfg.n_bins if cfg.flag_use_sensor else 0), HRL_ANT_MJ: 29, HRL_ANT_MAZE_MJ: 30 + 3 * cfg.n_bins,
            HRL_POINT_GATHER: 8 + food_obs_dim(cfg)}[k]


def act_dim(cfg):
    return 2 if cfg.env_kind == HRL_POINT_GATHER else 8


TWO_PI = 2 * math.pi
