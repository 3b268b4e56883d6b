"""The true-Bullet pinning kit must work the day pybullet is reachable: this test runs
`tools/dump_pybullet_truth.py` against a FAKE `pybullet` / `gym` / `hrl_pybullet_envs` stack whose "physics" is the
oracle, and feeds the file it writes to the consumers in tests/test_pybullet_truth.py.  It proves the fixture format
round-trips (state layout, joint order, key names, episode statistics) - not that the oracle equals Bullet: with the
oracle on both sides the comparison is exact by construction.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JOINTS = ["hip_1", "ankle_1", "hip_2", "ankle_2", "hip_3", "ankle_3", "hip_4", "ankle_4"]


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _FakeJoint:
    def __init__(self, env, k):
        self.env, self.k = env, k

    def get_state(self):
        f, _ = self.env.o.get_state()
        return float(f[0, 13 + self.k]), float(f[0, 21 + self.k])


class _FakeBullet:
    """The handful of pybullet calls the dump tool makes, answered from the oracle's state."""

    def __init__(self, env):
        self.env = env

    def getNumJoints(self, body):
        return 12

    def getBodyInfo(self, body):
        return (b"torso", b"ant")

    def getDynamicsInfo(self, body, link):
        mass = 65.44984694978736 if link == -1 else (13.51850726010076 if link in (2, 5, 8, 11) else 7.831583314284915)
        inertia = (2.7270769562411402,) * 3 if link == -1 else (0.1, 0.1, 0.17)
        return (mass, 1.5, inertia, (0, 0, 0), (0, 0, 0, 1))

    def getJointInfo(self, body, link):
        name = ("jointfix_%d" % link) if link % 3 == 0 else JOINTS[(link // 3) * 2 + (link % 3) - 1]
        return (link, name.encode(), 0, 0, 0, 0, 0.0, 0.0, -0.7, 0.7, 0, 0, ("link%d" % link).encode(), (0, 0, 1), (0, 0, 0), (0, 0, 0, 1), link - 1)

    def getPhysicsEngineParameters(self):
        return {"fixedTimeStep": 0.0165, "numSubSteps": 4, "numSolverIterations": 5, "contactERP": 0.9}

    def getBasePositionAndOrientation(self, body):
        f, _ = self.env.o.get_state()
        return tuple(f[0, 0:3]), tuple(f[0, 3:7])

    def getBaseVelocity(self, body):
        f, _ = self.env.o.get_state()
        return tuple(f[0, 7:10]), tuple(f[0, 10:13])


class _FakeEnv:
    """gym-style single env over a 1-env OracleVecEnv; exposes the attributes the dump tool reads."""

    def __init__(self, env_id):
        from oracle import oracle as O
        self.o = O.OracleVecEnv.make(env_id, 1, seed=0)
        self.o.cfg.auto_reset = 0
        self.action_space = types.SimpleNamespace(shape=(self.o.A,))
        self._p = _FakeBullet(self)
        self.robot = types.SimpleNamespace(objects=[0], jdict={n: _FakeJoint(self, k) for k, n in enumerate(JOINTS)})
        self._elapsed_steps = 0
        self._sync()

    def _sync(self):
        f, i = self.o.get_state()
        r = self.robot
        r.initial_z = float(f[0, 29]); self.potential = float(f[0, 30])
        r.walk_target_x, r.walk_target_y = float(f[0, 31]), float(f[0, 32])
        r.walk_target_dist = float(f[0, 33]); r.feet_contact = f[0, 34:38].copy()
        self._elapsed_steps = int(i[0, 0])
        self.goals = [0] * int(i[0, 3]); self.steps_since_goal_change = int(i[0, 4]); self._rewarded = bool(i[0, 5])

    def seed(self, s):
        return [s]

    def reset(self):
        obs = self.o.reset()
        self._sync()
        return obs[0]

    def step(self, a):
        obs, rew, done, info = self.o.step(np.asarray(a, np.float32)[None])
        self._sync()
        if self._elapsed_steps >= 60:   # short episodes keep the fake rollouts quick
            done[0] = True
        return obs[0], float(rew[0]), bool(done[0]), {}


def test_dump_tool_output_feeds_the_truth_tests(tmp_path, monkeypatch):
    monkeypatch.setitem(sys.modules, "pybullet", types.ModuleType("pybullet"))
    gym = types.ModuleType("gym")
    gym.make = lambda env_id: _FakeEnv(env_id)
    monkeypatch.setitem(sys.modules, "gym", gym)
    monkeypatch.setitem(sys.modules, "hrl_pybullet_envs", types.ModuleType("hrl_pybullet_envs"))
    tool = _load(os.path.join(ROOT, "tools", "dump_pybullet_truth.py"), "dump_pybullet_truth")
    env_id = "AntMazeBulletEnv-v0"
    tool.dump(env_id, str(tmp_path), n_states=12, n_episodes=2, seed=0)
    path = os.path.join(str(tmp_path), "pybullet_truth_%s.npz" % env_id)
    z = np.load(path, allow_pickle=True)
    assert z["state0_f"].shape == (12, 72) and z["state0_i"].shape == (12, 8) and z["action"].shape == (12, 8)
    assert z["obs"].shape == (12, 38) and len(z["episode_return"]) == 2 and (z["episode_length"] > 0).all()
    assert list(z["model_joint_names"][[1, 2, 4, 5]]) == ["hip_1", "ankle_1", "hip_2", "ankle_2"]
    # the consumers of tests/test_pybullet_truth.py accept the file (exact here: the fake stack IS the oracle)
    truth = _load(os.path.join(ROOT, "tests", "test_pybullet_truth.py"), "truth_tests")
    truth.test_model_constants_match_bullet(path)
    truth.test_oracle_one_step_vs_bullet(path)


def test_dump_tool_refuses_without_pybullet(capsys):
    tool = _load(os.path.join(ROOT, "tools", "dump_pybullet_truth.py"), "dump_pybullet_truth2")
    if importlib.util.find_spec("pybullet") is not None:
        pytest.skip("pybullet is installed here")
    old = sys.argv
    sys.argv = ["dump_pybullet_truth.py"]
    try:
        with pytest.raises(SystemExit) as e:
            tool.main()
    finally:
        sys.argv = old
    assert "pybullet is not installed" in str(e.value)
