#!/usr/bin/env python
"""Generate golden vectors by executing the REFERENCE's own pure-Python task logic.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The third-party modules the reference imports (pybullet, pybullet_envs, pybulletgym, gym)
are not installed, so ``_ref_stubs`` injects empty stand-ins and the reference's methods
are called unbound on ``SimpleNamespace`` stub objects that supply poses / robot state.
Fixtures contain inputs we chose and OUTPUTS of the reference code, never its source.
All float inputs are float32-representable so that the f32 CUDA path and the f64 oracle
see bit-identical inputs.
"""
import json
import math
import os
import sys
from types import SimpleNamespace as NS

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_stubs  # noqa: E402

gym = _ref_stubs.install("/root/reference")

import hrl_pybullet_envs  # noqa: E402  (runs the 5 register() calls)
from hrl_pybullet_envs.envs import intersection_utils as iu  # noqa: E402
from hrl_pybullet_envs.envs.gather.ant_gather_env import AntGatherBulletEnv  # noqa: E402
from hrl_pybullet_envs.envs.gather.gather_base import GatherBulletEnv  # noqa: E402
from hrl_pybullet_envs.envs.gather.gather_scene import GatherScene  # noqa: E402
from hrl_pybullet_envs.envs.gather.point_bot import PointBot  # noqa: E402
from hrl_pybullet_envs.envs.sizeable_enclosed_scene import SizeableEnclosedScene  # noqa: E402
from hrl_pybullet_envs.envs.ant_maze.maze_scene import MazeScene  # noqa: E402
from hrl_pybullet_envs.envs.ant_maze.ant_maze_bullet_env import AntMazeBulletEnv  # noqa: E402
from hrl_pybullet_envs.envs.ant_maze import ant_maze_bullet_env as maze_mod  # noqa: E402
from hrl_pybullet_envs.envs.ant_maze import ant_maze_mj_env as mazemj_mod  # noqa: E402
from hrl_pybullet_envs.envs.ant_flagrun.ant_flagrun_env import AntFlagrunBulletEnv  # noqa: E402
from hrl_pybullet_envs.envs.MjAnt import MjAnt, AntMjEnv  # noqa: E402
from hrl_pybullet_envs.utils import PositionEncoding  # noqa: E402


def f32(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


class ListRNG:
    """RandomState stand-in that replays a fixed list of uniforms (and counts draws)."""

    def __init__(self, uniforms):
        self.u = list(map(float, uniforms))
        self.i = 0

    def rand(self, n):
        out = np.array(self.u[self.i:self.i + n], dtype=np.float64)
        assert len(out) == n, "ListRNG exhausted"
        self.i += n
        return out


def pose_ns(xyz, rpy):
    return NS(xyz=lambda: np.array(xyz, dtype=np.float64), rpy=lambda: np.array(rpy, dtype=np.float64))


# --------------------------------------------------------------------------- 1. geometry
def gen_intersection(rng):
    pts = f32(rng.integers(-6, 7, size=(400, 8)) * 0.5)  # coarse grid => many colinear cases
    pts2 = f32(rng.uniform(-6, 6, size=(400, 8)))
    allp = np.concatenate([pts, pts2])
    seg = np.array([iu.segment_intersection(iu.Point(*p[0:2]), iu.Point(*p[2:4]),
                                            iu.Point(*p[4:6]), iu.Point(*p[6:8])) for p in allp])
    inter = []
    has = []
    for p in allp:
        r = iu.inf_intersection(iu.Point(*p[0:2]), iu.Point(*p[2:4]), iu.Point(*p[4:6]), iu.Point(*p[6:8]))
        has.append(r is not None)
        inter.append([r.x, r.y] if r is not None else [0.0, 0.0])
    qp = np.concatenate([f32(rng.uniform(-1, 1, size=(50, 2))),
                         np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1], [0.0, -0.0]], dtype=np.float64)])
    quad = np.array([iu.quadrant(iu.Point(*p)) for p in qp])
    np.savez_compressed(os.path.join(HERE, "intersection.npz"), seg_in=allp, seg_out=seg, inf_has=np.array(has),
             inf_xy=np.array(inter), quad_in=qp, quad_out=quad)


# --------------------------------------------------------------------------- 2. gather sensor
def ref_gather_sensor(cls, xy, yaw, objs, n_bins, sensor_range=20.0, span=math.pi):
    food = {i: [float(objs[i, 0]), float(objs[i, 1]), 0.1] for i in range(8)}
    poison = {i: [float(objs[i, 0]), float(objs[i, 1]), 0.1] for i in range(8, 16)}
    scene = NS(food=food, poison=poison, all_items={**food, **poison})
    torso = NS(get_pose=lambda: [xy[0], xy[1], 0.5, 0, 0, 0, 1], get_position=lambda: [xy[0], xy[1], 0.5],
               pose=lambda: pose_ns([xy[0], xy[1], 0.5], [0.0, 0.0, yaw]))
    stub = NS(n_bins=n_bins, sensor_span=span, sensor_range=sensor_range, debug=False,
              FOOD="food", POISON="poison", stadium_scene=scene, parts={"torso": torso},
              robot=NS(robot_body=torso),
              robot_body=NS(pose=lambda: pose_ns([xy[0], xy[1], 0.5], [0.0, 0.0, yaw])))
    stub.sq_dist_robot = lambda pos: cls.sq_dist_robot(stub, pos)
    dists = {i: stub.sq_dist_robot(p) for i, p in scene.all_items.items()}
    fr, pr = cls.get_sensor_readings(stub, dists)
    return fr, pr, np.array([dists[i] for i in range(16)])


def gen_gather_sensor(rng):
    M = 600
    xy = f32(rng.uniform(-7, 7, size=(M, 2)))
    yaw = f32(rng.uniform(-math.pi, math.pi, size=M))
    objs = f32(rng.uniform(-7, 7, size=(M, 16, 2)))
    # make a third of the cases dense around the robot so that bins collide / pickups happen
    objs[: M // 3] = f32(xy[: M // 3, None, :] + rng.uniform(-3, 3, size=(M // 3, 16, 2)))
    # hand cases: SURVEY 8c(1)
    xy[0] = f32([0.3, -0.2]); yaw[0] = f32(0.4)
    objs[0] = 50.0
    objs[0, 0] = [1.0, 0.5]; objs[0, 1] = [-2.0, 1.0]; objs[0, 8] = [0.5, 2.0]
    out = {}
    for n_bins, cls, tag in [(10, AntGatherBulletEnv, "ant"), (5, GatherBulletEnv, "point")]:
        F, P, D = [], [], []
        for m in range(M):
            fr, pr, d = ref_gather_sensor(cls, xy[m], float(yaw[m]), objs[m], n_bins)
            F.append(fr); P.append(pr); D.append(d)
        out[f"food_{tag}"] = np.array(F); out[f"poison_{tag}"] = np.array(P); out[f"d2_{tag}"] = np.array(D)
    np.savez_compressed(os.path.join(HERE, "gather_sensor.npz"), xy=xy, yaw=yaw, objs=objs, **out)


# --------------------------------------------------------------------------- 3. wall lidar
def gen_sense_walls(rng):
    M = 400
    maze = MazeScene(None, 9.8, 0.0165 / 4, 4)
    flag = SizeableEnclosedScene(None, 9.8, 0.0165 / 4, 4, (12, 12))
    bounds = {"maze": [[a.x, a.y, b.x, b.y] for a, b in maze.bounds],
              "flagrun": [[a.x, a.y, b.x, b.y] for a, b in flag.bounds]}
    res = {}
    for tag, scene, lo, hi in [("maze", maze, (-5, -9), (5, 9)), ("flagrun", flag, (-6, -6), (6, 6))]:
        xy = f32(rng.uniform(lo, hi, size=(M, 2)))
        yaw = f32(rng.uniform(-math.pi, math.pi, size=M))
        if tag == "maze":
            xy[0] = [-2, -5]; yaw[0] = 0.0
            yaw[1:20] = f32(np.round(rng.uniform(-2, 2, size=19)) * (math.pi / 2))  # axis-aligned rays
        full = np.array([scene.sense_walls(10, 2 * np.pi, 5.0, xy[m], float(yaw[m])) for m in range(M)])
        half = np.array([scene.sense_walls(8, np.pi, 4.0, xy[m], float(yaw[m])) for m in range(M)])
        res[f"{tag}_xy"] = xy; res[f"{tag}_yaw"] = yaw
        res[f"{tag}_full10_r5"] = full; res[f"{tag}_pi8_r4"] = half
        res[f"{tag}_bounds"] = np.array(bounds[tag], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "sense_walls.npz"), **res)


# --------------------------------------------------------------------------- 4. maze goal obs
def gen_maze_target(rng):
    M = 300
    maze = MazeScene(None, 9.8, 0.0165 / 4, 4)
    xy = f32(rng.uniform((-5, -9), (5, 9), size=(M, 2)))
    yaw = f32(rng.uniform(-math.pi, math.pi, size=M))
    targets = np.array(maze_mod._targets, dtype=np.float64)
    tid = rng.integers(0, 4, size=M)
    wtd = f32(rng.uniform(0.2, 8.0, size=M))  # walk_target_dist is an independent (Q1-distorted) input
    xy[0] = [2, 1]; yaw[0] = f32(0.3); tid[0] = 2; wtd[0] = 2.0       # bin 7 = 0.6
    xy[1] = [-2, -5]; yaw[1] = 0.0; tid[1] = 3; wtd[1] = 2.0           # occluded by the box line y=-2
    vec0, vec1, sens = [], [], []
    for m in range(M):
        stub = NS(target=targets[tid[m]], n_bins=10, sensor_span=2 * np.pi, sensor_range=5.0, debug=0,
                  robot_body=NS(pose=lambda m=m: pose_ns([xy[m, 0], xy[m, 1], 0.5], [0, 0, float(yaw[m])])),
                  robot=NS(walk_target_dist=float(wtd[m])), scene=maze)
        stub.target_encoding = PositionEncoding.normed_vec
        vec0.append(np.asarray(AntMazeBulletEnv.get_target_vec_obs(stub), dtype=np.float64))
        stub.target_encoding = PositionEncoding.angle
        vec1.append(np.asarray(AntMazeBulletEnv.get_target_vec_obs(stub), dtype=np.float64))
        sens.append(AntMazeBulletEnv.get_target_sensor_obs(stub))
    np.savez_compressed(os.path.join(HERE, "maze_target.npz"), xy=xy, yaw=yaw, tid=tid, wtd=wtd, targets=targets,
             targets_mj=np.array(mazemj_mod._targets, dtype=np.float64),
             vec_normed=np.array(vec0), vec_angle=np.array(vec1), sensor=np.array(sens))


# --------------------------------------------------------------------------- 5. placement RNG logic
def gen_random_on_plane(rng):
    fake = _ref_stubs.FakeBullet()
    sc = GatherScene(None, 9.8, 0.0165 / 4, 4, (15, 15), 8, 8, 2.0, True)
    # (a) _random_on_plane with replayed uniforms, several avoid points
    M = 200
    u = f32(((rng.integers(0, 1 << 24, size=(M, 64))).astype(np.float64)) / (1 << 24))
    avoid = f32(rng.uniform(-7, 7, size=(M, 2)))
    u[:, 0:2] = f32((avoid + 7.0) / 14.0)  # force at least one rejection per case
    pos, used = [], []
    for m in range(M):
        sc.rs = ListRNG(u[m])
        p = sc._random_on_plane(list(avoid[m]))
        pos.append(p); used.append(sc.rs.i)
    # (b) a whole episode_restart with replayed uniforms -> 16 positions, in food-then-poison order
    u2 = f32(((rng.integers(0, 1 << 24, size=400)).astype(np.float64)) / (1 << 24))
    sc.rs = ListRNG(u2)
    sc.episode_restart(fake)
    first = np.array(list(sc.food.values()) + list(sc.poison.values()))
    used_first = sc.rs.i
    sc.episode_restart(fake)  # second restart: only the re-randomise pass draws
    second = np.array(list(sc.food.values()) + list(sc.poison.values()))
    used_second = sc.rs.i
    # (c) reward_collision semantics
    ids_food = list(sc.food.keys()); ids_poison = list(sc.poison.keys())
    rew = [sc.reward_collision(ids_food[0], [0.0, 0.0, 0.5]), sc.reward_collision(ids_poison[0], [0.0, 0.0, 0.5]),
           sc.reward_collision(9999, [0.0, 0.0, 0.5])]
    sc.respawn = False
    rew_norespawn = sc.reward_collision(ids_food[1], [0.0, 0.0, 0.5])
    parked = sc.food[ids_food[1]]
    np.savez_compressed(os.path.join(HERE, "random_on_plane.npz"), u=u, avoid=avoid, pos=np.array(pos), used=np.array(used),
             u2=u2, restart1=first, used1=used_first, restart2=second, used2=used_second,
             rew=np.array(rew), rew_norespawn=rew_norespawn, parked=np.array(parked, dtype=np.float64),
             n_loaded=fake.next_id)


# --------------------------------------------------------------------------- 6. flagrun goals
def gen_flagrun_goals():
    stub = NS(size=10, mpi_common_rand=np.random.RandomState(123))
    throwaway = AntFlagrunBulletEnv.create_target(stub)
    stub.create_target = lambda: AntFlagrunBulletEnv.create_target(stub)
    AntFlagrunBulletEnv.create_targets(stub, 100)
    g1 = np.array(stub.goals)
    AntFlagrunBulletEnv.create_targets(stub, 100)
    g2 = np.array(stub.goals)
    # rejection rule with replayed uniforms: RandomState.uniform(lo,hi) = lo + (hi-lo)*u
    np.savez_compressed(os.path.join(HERE, "flagrun_goals.npz"), throwaway=np.array(throwaway), goals1=g1, goals2=g2)


# --------------------------------------------------------------------------- 7. AntGather.step task layer
GATHER_STEP_CASES = [("ant", AntGatherBulletEnv, 10, 28), ("point", GatherBulletEnv, 5, 8),
                     ("antabs", AntGatherBulletEnv, 5, 28)]  # antabs: use_sensor=False -> get_abs_pos


def gen_gather_step(rng, cases=GATHER_STEP_CASES, out="gather_step.npz"):
    """Whole `AntGatherBulletEnv.step` / `GatherBulletEnv.step` with a stub robot: given the
    post-physics robot state (calc_state output, torso pose) and the item table, the reference
    computes pickups, respawns, the sensor, alive/done and the reward."""
    M = 400
    fake = _ref_stubs.FakeBullet()
    res = {}
    for tag, cls, n_bins, sdim in cases:
        state = f32(rng.uniform(-1, 1, size=(M, sdim)))
        xyz = f32(np.concatenate([rng.uniform(-7, 7, size=(M, 2)), rng.uniform(0.15, 0.9, size=(M, 1))], axis=1))
        rpy = f32(rng.uniform(-math.pi, math.pi, size=(M, 3)) * [0.2, 0.2, 1.0])
        objs = f32(xyz[:, None, :2] + rng.uniform(-4, 4, size=(M, 16, 2)))
        objs[M // 2:] = f32(rng.uniform(-7, 7, size=(M - M // 2, 16, 2)))
        u = f32(((rng.integers(0, 1 << 24, size=(M, 256))).astype(np.float64)) / (1 << 24))
        state[5, 3] = np.inf  # non-finite guard
        state[6, 0] = np.nan
        OBS, REW, DONE, FR, DR, NEWO, USED = [], [], [], [], [], [], []
        for m in range(M):
            sc = GatherScene(None, 9.8, 0.0165 / 4, 4, (15, 15), 8, 8, 2.0, True)
            sc._p = fake
            sc.food = {i: [float(objs[m, i, 0]), float(objs[m, i, 1]), 0.1] for i in range(8)}
            sc.poison = {i: [float(objs[m, i, 0]), float(objs[m, i, 1]), 0.1] for i in range(8, 16)}
            sc.rs = ListRNG(u[m])
            initial_z = 0.75 if tag.startswith("ant") else 1.0
            st = state[m].copy()
            st[0] = xyz[m, 2] - initial_z
            torso = NS(get_pose=lambda m=m: [*xyz[m], 0, 0, 0, 1], get_position=lambda m=m: list(xyz[m]),
                       pose=lambda m=m: pose_ns(xyz[m], rpy[m]))
            if tag.startswith("ant"):
                alive = lambda z, pitch: +1 if z > 0.26 else -1
            else:
                alive = lambda z, pitch: PointBot.alive_bonus(None, z, pitch)
            robot = NS(apply_action=lambda a: None, calc_state=lambda st=st: st.astype(np.float32),
                       alive_bonus=alive, initial_z=initial_z, body_rpy=rpy[m], robot_body=torso, objects=[0])
            stub = NS(robot=robot, scene=NS(global_step=lambda: None), stadium_scene=sc, parts={"torso": torso},
                      robot_body=torso, robot_coll_dist=1, n_bins=n_bins, sensor_span=np.pi, sensor_range=20.0,
                      use_sensor=not tag.endswith("abs"), dying_cost=-10, debug=False, FOOD="food", POISON="poison", _p=fake)
            stub.sq_dist_robot = lambda pos, stub=stub: cls.sq_dist_robot(stub, pos)
            stub.get_food_obs = lambda d, stub=stub: cls.get_food_obs(stub, d)
            stub.get_sensor_readings = lambda d, stub=stub: cls.get_sensor_readings(stub, d)
            stub.get_abs_pos = lambda d, stub=stub: cls.get_abs_pos(stub, d)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                obs, rew, done, info = cls.step(stub, np.zeros(8 if tag.startswith("ant") else 2))
            OBS.append(obs); REW.append(rew); DONE.append(done)
            FR.append(info["food_rew"]); DR.append(info["dead_rew"])
            NEWO.append(np.array([sc.all_items[i][:2] for i in range(16)]))
            USED.append(sc.rs.i)
        res.update({f"{tag}_state": state, f"{tag}_xyz": xyz, f"{tag}_rpy": rpy, f"{tag}_objs": objs, f"{tag}_u": u,
                    f"{tag}_obs": np.array(OBS, dtype=np.float64), f"{tag}_rew": np.array(REW, dtype=np.float64),
                    f"{tag}_done": np.array(DONE), f"{tag}_food_rew": np.array(FR, dtype=np.float64),
                    f"{tag}_dead_rew": np.array(DR, dtype=np.float64), f"{tag}_new_objs": np.array(NEWO),
                    f"{tag}_used": np.array(USED)})
    np.savez_compressed(os.path.join(HERE, out), **res)


# --------------------------------------------------------------------------- 8. AntMaze.step task layer
def gen_maze_step(rng):
    M = 300
    maze = MazeScene(None, 9.8, 0.0165 / 4, 4)
    ant_obs = f32(rng.uniform(-1, 1, size=(M, 28)))
    xy = f32(rng.uniform((-5, -9), (5, 9), size=(M, 2)))
    yaw = f32(rng.uniform(-math.pi, math.pi, size=M))
    inner_rew = f32(rng.uniform(-2, 2, size=M))
    inner_done = rng.uniform(size=M) < 0.2
    wtd = f32(rng.uniform(0.5, 4.0, size=M))
    tid = rng.integers(0, 4, size=M)
    targets = np.array(maze_mod._targets, dtype=np.float64)
    t_before = rng.integers(0, 9, size=M)  # used by the max_steps variant only
    # default kwargs first, then the non-default ones (ant_maze_bullet_env.py:23-25)
    variants = {"": {}, "_sense_target": dict(sense_target=True),
                "_max_steps": dict(max_steps=7, done_at_target=False),
                "_targ_dist": dict(targ_dist_rew=True, inner_rew_weight=0.5),
                "_angle_nowalls": dict(target_encoding=PositionEncoding.angle, sense_walls=False)}
    res = {}
    for tag, kw in variants.items():
        OBS, REW, DONE = [], [], []
        for m in range(M):
            # bind a fake inner step onto the stub base class the reference env derives from
            maze_mod.AntBulletEnv.step = lambda self, a, m=m: (ant_obs[m].astype(np.float32), float(inner_rew[m]),
                                                              bool(inner_done[m]), {})
            env = AntMazeBulletEnv.__new__(AntMazeBulletEnv)
            env.t = int(t_before[m]) if tag == "_max_steps" else 0
            env.debug = 0; env.target = targets[tid[m]]; env.inner_rew_weight = 0; env.tol = 1.5
            env.done_at_target = True; env.max_steps = -1; env.targ_dist_rew = False
            env.sense_walls = True; env.sense_target = False; env.n_bins = 10; env.sensor_span = 2 * np.pi
            env.sensor_range = 5.0; env.target_encoding = PositionEncoding.normed_vec
            for k, v in kw.items():
                setattr(env, k, v)
            env.scene = maze
            env.robot = NS(walk_target_dist=float(wtd[m]), body_real_xyz=np.array([xy[m, 0], xy[m, 1], 0.5]))
            env.robot_body = NS(pose=lambda m=m: pose_ns([xy[m, 0], xy[m, 1], 0.5], [0, 0, float(yaw[m])]))
            obs, rew, done, info = env.step(np.zeros(8))
            OBS.append(obs); REW.append(rew); DONE.append(done)
        res.update({"obs" + tag: np.array(OBS, dtype=np.float64), "rew" + tag: np.array(REW, dtype=np.float64),
                    "done" + tag: np.array(DONE)})
    res["t_before_max_steps"] = t_before
    del maze_mod.AntBulletEnv.step
    np.savez_compressed(os.path.join(HERE, "maze_step.npz"), ant_obs=ant_obs, xy=xy, yaw=yaw, inner_rew=inner_rew,
             inner_done=inner_done, wtd=wtd, tid=tid, targets=targets, **res)


# --------------------------------------------------------------------------- 8b. AntMazeMj.step task layer
def gen_maze_mj_step(rng):
    """Whole `AntMazeMjEnv.step` / `_get_obs` (ant_maze_mj_env.py:57-79) with a stub inner `AntMjEnv.step`: 60-d
    observation [mj 29 | walls 10 | 0 x 10 | 0 x 10 | t * 0.001], reward inner * weight (+1 inside tol), done."""
    from hrl_pybullet_envs.envs.ant_maze.ant_maze_mj_env import AntMazeMjEnv
    M = 300
    maze = MazeScene(None, 9.8, 0.0165 / 4, 4)
    mj_obs = f32(rng.uniform(-1, 1, size=(M, 29)))
    mj_obs[:, 0:2] = f32(rng.uniform((-5, -9), (5, 9), size=(M, 2)))   # the lidar is cast from ant_obs[:2] (:58-59)
    yaw = f32(rng.uniform(-math.pi, math.pi, size=M))
    inner_rew = f32(rng.uniform(-2, 2, size=M))
    inner_done = rng.uniform(size=M) < 0.2
    wtd = f32(rng.uniform(0.5, 4.0, size=M))
    t_before = rng.integers(0, 2000, size=M)
    variants = {"": {}, "_weight": dict(inner_rew_weight=0.5, tol=2.5)}
    res = {}
    for tag, kw in variants.items():
        OBS, REW, DONE, T_AFTER = [], [], [], []
        for m in range(M):
            mazemj_mod.AntMjEnv.step = lambda self, a, m=m: (mj_obs[m].copy(), float(inner_rew[m]), bool(inner_done[m]), {})
            env = AntMazeMjEnv.__new__(AntMazeMjEnv)
            env.t = int(t_before[m]); env.debug = 0; env.inner_rew_weight = 0; env.tol = 1.5
            env.n_bins = 10; env.sensor_span = 2 * np.pi; env.sensor_range = 5.0
            for k, v in kw.items():
                setattr(env, k, v)
            env.scene = maze
            env.robot = NS(walk_target_dist=float(wtd[m]))
            env.robot_body = NS(pose=lambda m=m: pose_ns([mj_obs[m, 0], mj_obs[m, 1], 0.5], [0, 0, float(yaw[m])]))
            obs, rew, done, info = env.step(np.zeros(8))
            OBS.append(obs); REW.append(rew); DONE.append(done); T_AFTER.append(env.t)
        res.update({"obs" + tag: np.array(OBS, dtype=np.float64), "rew" + tag: np.array(REW, dtype=np.float64),
                    "done" + tag: np.array(DONE), "t_after" + tag: np.array(T_AFTER)})
    del mazemj_mod.AntMjEnv.step
    np.savez_compressed(os.path.join(HERE, "maze_mj_step.npz"), mj_obs=mj_obs, yaw=yaw, inner_rew=inner_rew, inner_done=inner_done,
                        wtd=wtd, t_before=t_before, targets=np.array(mazemj_mod._targets, dtype=np.float64), **res)


# --------------------------------------------------------------------------- 9. Flagrun.step sequence
def gen_flagrun_step(rng):
    """Drive the reference's Flagrun step logic with a scripted sequence of walk_target_dist
    values and inner (reward, done) pairs; record reward, done, the target after the step,
    steps_since_goal_change and the rewarded flag."""
    import hrl_pybullet_envs.envs.ant_flagrun.ant_flagrun_env as fr_mod
    T = 700
    wtd = f32(rng.uniform(0.1, 3.0, size=T))
    wtd[rng.uniform(size=T) < 0.85] += 1.0  # mostly far from the goal
    inner_r = f32(rng.uniform(-1, 1, size=T))
    seqs = {}
    for tag, n_goals, timeout, switch in [("a", 100, 200, True), ("b", 3, 50, True), ("c", 5, 40, False)]:
        env = AntFlagrunBulletEnv.__new__(AntFlagrunBulletEnv)
        env.size = 10; env.tol = 0.5; env.max_targets = n_goals; env.max_target_dist = 0; env.timeout = timeout
        env.switch_flag_on_collision = switch; env.debug = False; env.use_sensor = False; env.isRender = False
        env.flag = None
        env.mpi_common_rand = np.random.RandomState(123)
        env.create_target()
        env.steps_since_goal_change = 0; env.goals = []; env._rewarded = False
        env._sq_dist_goal = 0; env._goal_start_pos = np.array([0, 0])
        pos = np.array([0.0, 0.0, 0.5])
        body = NS(get_position=lambda: pos)
        env.robot = NS(walk_target_dist=1.0, body_real_xyz=pos, robot_body=body, walk_target_x=0, walk_target_y=0,
                       calc_potential=lambda: -1.0, calc_state=lambda: np.zeros(28, dtype=np.float32))
        env.create_targets(n_goals)
        goals0 = np.array(env.goals)
        env.next_target()
        first_target = np.array(env.goal)
        t = {"i": 0}
        fr_mod.AntBulletEnv.step = lambda self, a: (np.ones(28, dtype=np.float32), float(inner_r[t["i"]]), False, {})
        R, D, TG, SS, RW = [], [], [], [], []
        for i in range(T):
            t["i"] = i
            env.robot.walk_target_dist = float(wtd[i])
            s, r, d, info = env.step(np.zeros(8))
            R.append(r); D.append(d); TG.append(list(env.goal)); SS.append(env.steps_since_goal_change)
            RW.append(env._rewarded)
            if d:
                break
        del fr_mod.AntBulletEnv.step
        seqs.update({f"{tag}_goals0": goals0, f"{tag}_first_target": first_target, f"{tag}_rew": np.array(R),
                     f"{tag}_done": np.array(D), f"{tag}_target": np.array(TG), f"{tag}_since": np.array(SS),
                     f"{tag}_rewarded": np.array(RW), f"{tag}_n_goals": n_goals, f"{tag}_timeout": timeout,
                     f"{tag}_switch": int(switch)})
    np.savez_compressed(os.path.join(HERE, "flagrun_step.npz"), wtd=wtd, inner_r=inner_r, **seqs)


# --------------------------------------------------------------------------- 10. PointBot + MjAnt
def gen_robots(rng):
    M = 200
    xyz = f32(rng.uniform(-7, 7, size=(M, 3)))
    rpy = f32(rng.uniform(-math.pi, math.pi, size=(M, 3)))
    vel = f32(rng.uniform(-3, 3, size=(M, 3)))
    S = []
    for m in range(M):
        stub = NS(robot_body=NS(pose=lambda m=m: pose_ns(xyz[m], rpy[m]), speed=lambda m=m: vel[m]),
                  walk_target_x=0, walk_target_y=0, initial_z=1)
        S.append(PointBot.calc_state(stub))
    act = f32(rng.uniform(-1, 1, size=(M, 2)))
    forces = []
    for m in range(M):
        rec = {}
        p = NS(getBasePositionAndOrientation=lambda o: ((0, 0, 0), (0, 0, 0, 1)), WORLD_FRAME=1,
               applyExternalForce=lambda o, l, f, pos, fl: rec.setdefault("f", list(f)))
        PointBot.apply_action(NS(_p=p, objects=[0]), act[m])
        forces.append(rec["f"])
    # AntMjEnv.step reward composition with a stub MjAnt
    st = f32(rng.uniform(-1, 1, size=(M, 29)))
    st[:, 2] = f32(rng.uniform(0.2, 0.9, size=M))
    pot_old = f32(rng.uniform(-60, 0, size=M)); pot_new = f32(pot_old + rng.uniform(-1, 1, size=M))
    jal = rng.integers(0, 9, size=M)
    REW, DONE = [], []
    for m in range(M):
        robot = NS(apply_action=lambda a: None, calc_state=lambda m=m: st[m], initial_z=0.75,
                   body_rpy=np.zeros(3), calc_potential=lambda m=m: float(pot_new[m]), feet=[], feet_contact=np.zeros(4),
                   joints_at_limit=int(jal[m]))
        robot.alive_bonus = lambda z, pitch, robot=robot: MjAnt.alive_bonus(robot, z, pitch)
        env = NS(robot=robot, scene=NS(global_step=lambda: None), potential=float(pot_old[m]), ground_ids=set(),
                 joints_at_limit_cost=-0.1, HUD=lambda *a: None, reward=0)
        s, r, d, _ = AntMjEnv.step(env, np.zeros(8))
        REW.append(r); DONE.append(d)
    np.savez_compressed(os.path.join(HERE, "robots.npz"), xyz=xyz, rpy=rpy, vel=vel, point_state=np.array(S, dtype=np.float64),
             act=act, force=np.array(forces, dtype=np.float64), mj_state=st, pot_old=pot_old, pot_new=pot_new,
             jal=jal, mj_rew=np.array(REW, dtype=np.float64), mj_done=np.array(DONE))


def gen_registry():
    reg = [{"id": r["id"], "max_episode_steps": r["max_episode_steps"],
            "entry_point": r["entry_point"].split(":")[1]} for r in gym.registered]
    with open(os.path.join(HERE, "registry.json"), "w") as f:
        json.dump(reg, f, indent=1)


if __name__ == "__main__":
    rng = np.random.default_rng(20261018)
    gen_intersection(rng)
    gen_gather_sensor(rng)
    gen_sense_walls(rng)
    gen_maze_target(rng)
    gen_random_on_plane(rng)
    gen_flagrun_goals()
    gen_gather_step(rng)
    # added later, with its own generator so that the files above stay byte-identical: PointGather use_sensor=False
    # (GatherBulletEnv.get_abs_pos, gather_base.py:170-187), 4 nearest food + 4 nearest poison items
    gen_gather_step(np.random.default_rng(20261019), cases=[("pointabs", GatherBulletEnv, 4, 8)], out="gather_step_pointabs.npz")
    gen_maze_step(rng)
    gen_flagrun_step(rng)
    gen_robots(rng)
    gen_registry()
    gen_maze_mj_step(np.random.default_rng(20261020))   # added in round 2 with its own generator (files above unchanged)
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith((".npz", ".json")):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))
