#!/usr/bin/env python
"""dump_pybullet_truth.py - close "parity unpinned" (SURVEY.md section 8c / 8f item 1).

Run this on ANY machine that has the real stack (``pip install pybullet gym==0.21 hrl_pybullet_envs``;
the build container and the GPU boxes do not).  It drives the UNMODIFIED
reference envs and writes, per env id, ``tests/golden/pybullet_truth_<id>.npz`` holding

  * the model constants Bullet really uses (link masses, local inertia diagonals, inertial frame
    offsets, joint axes / limits / damping, contact filter flags)  -> checks SURVEY.md App. A.3 / C.1
  * K saved states in THIS repo's checkpoint layout (include/hrl_b200.h HRL_SF_* / HRL_SI_*), the
    action applied, and the reference's one-step result: next state, observation, reward, done
  * fixed-seed random-action rollouts: per-episode return and length               -> statistics

``tests/test_pybullet_truth.py`` picks the files up automatically: once they are committed the
oracle and the CUDA path are checked against real Bullet with the north-star tolerances
(1e-3 m, 1e-2 rad/s, 1e-4 reward) and nothing else in the repo has to change.

    python tools/dump_pybullet_truth.py [--out tests/golden] [--states 256] [--episodes 32]

Nothing in the product imports this file.
"""
import argparse
import os
import sys

import numpy as np

IDS = ["AntGatherBulletEnv-v0", "AntMazeBulletEnv-v0", "AntFlagrunBulletEnv-v0", "PointGatherBulletEnv-v0",
       "AntMazeMjEnv-v0"]
JOINT_ORDER = ["hip_1", "ankle_1", "hip_2", "ankle_2", "hip_3", "ankle_3", "hip_4", "ankle_4"]
# offsets of include/hrl_b200.h
SF_POS, SF_QUAT, SF_LINVEL, SF_ANGVEL, SF_Q, SF_QD = 0, 3, 7, 10, 13, 21
SF_INITIAL_Z, SF_POTENTIAL, SF_TARGET, SF_WTD, SF_FEET, SF_ITEMS = 29, 30, 31, 33, 34, 38
STATE_F, STATE_I = 72, 8


def unwrap(env):
    while hasattr(env, "env"):
        env = env.env
    return env


def model_constants(env):
    """getDynamicsInfo / getJointInfo of every link of the robot body."""
    e = unwrap(env)
    p = e._p
    body = e.robot.objects[0] if isinstance(e.robot.objects, (list, tuple)) else e.robot.objects
    out = {"link_names": [], "mass": [], "inertia_diag": [], "inertial_pos": [], "inertial_orn": [], "lateral_friction": [],
           "joint_names": [], "joint_type": [], "joint_axis": [], "joint_lower": [], "joint_upper": [], "joint_damping": [],
           "joint_friction": [], "parent_frame_pos": [], "parent_frame_orn": [], "parent_index": []}
    for link in range(-1, p.getNumJoints(body)):
        d = p.getDynamicsInfo(body, link)
        out["mass"].append(d[0]); out["lateral_friction"].append(d[1]); out["inertia_diag"].append(d[2])
        out["inertial_pos"].append(d[3]); out["inertial_orn"].append(d[4])
        if link == -1:
            out["link_names"].append(p.getBodyInfo(body)[0].decode())
            continue
        j = p.getJointInfo(body, link)
        out["joint_names"].append(j[1].decode()); out["joint_type"].append(j[2]); out["joint_damping"].append(j[6])
        out["joint_friction"].append(j[7]); out["joint_lower"].append(j[8]); out["joint_upper"].append(j[9])
        out["link_names"].append(j[12].decode()); out["joint_axis"].append(j[13]); out["parent_frame_pos"].append(j[14])
        out["parent_frame_orn"].append(j[15]); out["parent_index"].append(j[16])
    pe = p.getPhysicsEngineParameters()
    out["engine"] = np.array([pe.get("fixedTimeStep", np.nan), pe.get("numSubSteps", np.nan), pe.get("numSolverIterations", np.nan),
                              pe.get("erp", np.nan), pe.get("contactERP", np.nan), pe.get("frictionERP", np.nan),
                              pe.get("contactBreakingThreshold", np.nan), pe.get("gravityAccelerationZ", np.nan)], float)
    return {k: np.array(v) for k, v in out.items()}


def state_of(env):
    """The env's state in this repo's checkpoint layout (one row)."""
    e = unwrap(env)
    p = e._p
    f = np.zeros(STATE_F, np.float64); iv = np.zeros(STATE_I, np.int32)
    r = e.robot
    body = r.objects[0] if isinstance(r.objects, (list, tuple)) else r.objects
    pos, orn = p.getBasePositionAndOrientation(body)
    lin, ang = p.getBaseVelocity(body)
    f[SF_POS:SF_POS + 3] = pos; f[SF_QUAT:SF_QUAT + 4] = orn; f[SF_LINVEL:SF_LINVEL + 3] = lin; f[SF_ANGVEL:SF_ANGVEL + 3] = ang
    if hasattr(r, "jdict"):
        for k, name in enumerate(JOINT_ORDER):
            if name in r.jdict:
                q, qd = r.jdict[name].get_state()
                f[SF_Q + k] = q; f[SF_QD + k] = qd
    f[SF_INITIAL_Z] = getattr(r, "initial_z", 0.0) or 0.0
    f[SF_POTENTIAL] = getattr(e, "potential", 0.0) or 0.0
    f[SF_TARGET] = getattr(r, "walk_target_x", 0.0); f[SF_TARGET + 1] = getattr(r, "walk_target_y", 0.0)
    f[SF_WTD] = getattr(r, "walk_target_dist", 0.0) or 0.0
    fc = getattr(r, "feet_contact", None)
    if fc is not None:
        f[SF_FEET:SF_FEET + len(fc)] = fc
    sc = getattr(e, "stadium_scene", None)
    if sc is not None and hasattr(sc, "food"):
        items = [v[:2] for v in sc.food.values()] + [[100.0, 0.0]] * (8 - len(sc.food)) + \
                [v[:2] for v in sc.poison.values()] + [[100.0, 0.0]] * (8 - len(sc.poison))
        f[SF_ITEMS:SF_ITEMS + 32] = np.array(items, float).reshape(-1)
    iv[0] = getattr(env, "_elapsed_steps", 0) or 0
    iv[3] = len(getattr(e, "goals", []) or [])
    iv[4] = getattr(e, "steps_since_goal_change", 0)
    iv[5] = int(bool(getattr(e, "_rewarded", False)))
    return f, iv


def dump(env_id, out_dir, n_states, n_episodes, seed):
    import gym
    import hrl_pybullet_envs  # noqa: F401  (registers the ids)
    env = gym.make(env_id)
    env.seed(seed)
    rng = np.random.RandomState(seed)
    env.reset()
    consts = model_constants(env)
    A = env.action_space.shape[0]
    S0, I0, ACT, S1, I1, OBS, REW, DONE = [], [], [], [], [], [], [], []
    every = 7  # sample a state every `every` steps so that flight, landing and walking phases are all covered
    t = 0
    while len(S0) < n_states:
        a = rng.uniform(-1, 1, A).astype(np.float32)
        take = (t % every == 0)
        if take:
            f, iv = state_of(env)
        obs, rew, done, info = env.step(a)
        if take:
            f1, iv1 = state_of(env)
            S0.append(f); I0.append(iv); ACT.append(a); S1.append(f1); I1.append(iv1)
            OBS.append(np.asarray(obs, np.float64)); REW.append(rew); DONE.append(done)
        t += 1
        if done:
            env.reset()
    returns, lengths = [], []
    for ep in range(n_episodes):
        env.reset(); R = 0.0; L = 0
        while True:
            obs, rew, done, info = env.step(rng.uniform(-1, 1, A).astype(np.float32))
            R += rew; L += 1
            if done:
                break
        returns.append(R); lengths.append(L)
    path = os.path.join(out_dir, "pybullet_truth_%s.npz" % env_id)
    np.savez_compressed(path, state0_f=np.array(S0), state0_i=np.array(I0), action=np.array(ACT), state1_f=np.array(S1),
                        state1_i=np.array(I1), obs=np.array(OBS), rew=np.array(REW), done=np.array(DONE),
                        episode_return=np.array(returns), episode_length=np.array(lengths), seed=seed,
                        **{"model_" + k: v for k, v in consts.items()})
    print("wrote", path, "states", len(S0), "episodes", n_episodes, "mean return %.2f" % float(np.mean(returns)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    ap.add_argument("--states", type=int, default=256)
    ap.add_argument("--episodes", type=int, default=32)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--ids", nargs="*", default=IDS)
    args = ap.parse_args()
    try:
        import pybullet  # noqa: F401
    except ImportError:
        sys.exit("pybullet is not installed here: run this on a machine with the reference stack (see the docstring)")
    for env_id in args.ids:
        dump(env_id, args.out, args.states, args.episodes, args.seed)


if __name__ == "__main__":
    main()
