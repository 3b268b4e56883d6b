"""Stub modules that let the reference's pure-Python task logic be imported in a
container without pybullet / pybullet_envs / pybulletgym / gym.

Only used by ``make_golden.py`` (run in the build container where /root/reference is
mounted).  Nothing here ships to the GPU box at test time - the generated ``*.npz`` /
``*.json`` fixtures do.
"""
import sys
import types
import hashlib

import numpy as np


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


class _Any:
    """Base class accepting any ctor args; used for every third-party base class."""

    def __init__(self, *a, **k):
        pass

    def episode_restart(self, *a, **k):
        pass


def legacy_np_random(seed=None):
    """gym<=0.21 ``seeding.np_random``: RandomState seeded from sha512(str(seed)).

    Restated from memory of gym 0.17-0.21; the exact stream does not
    matter for the fixtures because every fixture that consumes randomness injects a
    recording RNG instead.
    """
    if seed is None:
        seed = 0
    h = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    words = [int.from_bytes(h[i:i + 4], "little") for i in range(0, 8, 4)]
    rs = np.random.RandomState()
    rs.seed(words)
    return rs, seed


class FakeBullet:
    """Counts loadURDF ids, remembers poses, ignores everything else."""

    WORLD_FRAME = 1

    def __init__(self):
        self.next_id = 0
        self.pose = {}

    def loadURDF(self, path, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), *a, **k):
        i = self.next_id
        self.next_id += 1
        self.pose[i] = (str(path).rsplit("/", 1)[-1], list(basePosition))
        return i

    def resetBasePositionAndOrientation(self, i, pos, orn):
        name = self.pose.get(i, ("?", None))[0]
        self.pose[i] = (name, list(pos))

    def changeDynamics(self, *a, **k):
        pass

    def configureDebugVisualizer(self, *a, **k):
        pass

    def addUserDebugLine(self, *a, **k):
        pass

    def addUserDebugText(self, *a, **k):
        pass


def install(reference_root="/root/reference"):
    pb = _mod("pybullet")
    pb.getQuaternionFromEuler = lambda e: (0.0, 0.0, 0.0, 1.0)
    pb.COV_ENABLE_PLANAR_REFLECTION = 0
    pb.WORLD_FRAME = 1
    _mod("pybullet_data").getDataPath = lambda: "/nonexistent"

    pe = _mod("pybullet_envs")
    gl = _mod("pybullet_envs.gym_locomotion_envs")
    eb = _mod("pybullet_envs.env_bases")
    rb = _mod("pybullet_envs.robot_bases")
    sa = _mod("pybullet_envs.scene_abstract")

    class AntBulletEnv(_Any):
        # the golden harness replaces `step`/`reset` per fixture
        pass

    class WalkerBaseBulletEnv(_Any):
        electricity_cost = -2.0
        stall_torque_cost = -0.1
        joints_at_limit_cost = -0.1

    gl.AntBulletEnv = AntBulletEnv
    gl.WalkerBaseBulletEnv = WalkerBaseBulletEnv
    eb.MJCFBaseBulletEnv = type("MJCFBaseBulletEnv", (_Any,), {})
    rb.MJCFBasedRobot = type("MJCFBasedRobot", (_Any,), {})
    rb.BodyPart = type("BodyPart", (_Any,), {})
    rb.Pose_Helper = type("Pose_Helper", (_Any,), {})
    sa.Scene = type("Scene", (_Any,), {})
    pe.gym_locomotion_envs, pe.env_bases, pe.robot_bases, pe.scene_abstract = gl, eb, rb, sa

    for name in ["pybulletgym", "pybulletgym.envs", "pybulletgym.envs.mujoco",
                 "pybulletgym.envs.mujoco.envs", "pybulletgym.envs.mujoco.envs.locomotion",
                 "pybulletgym.envs.mujoco.envs.locomotion.walker_base_env",
                 "pybulletgym.envs.mujoco.robots", "pybulletgym.envs.mujoco.robots.locomotors",
                 "pybulletgym.envs.mujoco.robots.locomotors.walker_base",
                 "pybulletgym.envs.mujoco.robots.robot_bases",
                 "pybulletgym.envs.mujoco.robots.locomotors.ant"]:
        _mod(name)
    sys.modules["pybulletgym.envs.mujoco.envs.locomotion.walker_base_env"].WalkerBaseMuJoCoEnv = \
        type("WalkerBaseMuJoCoEnv", (_Any,), {"joints_at_limit_cost": -0.1})
    sys.modules["pybulletgym.envs.mujoco.robots.locomotors.walker_base"].WalkerBase = \
        type("WalkerBase", (_Any,), {})
    sys.modules["pybulletgym.envs.mujoco.robots.robot_bases"].MJCFBasedRobot = \
        type("MJCFBasedRobot", (_Any,), {})
    sys.modules["pybulletgym.envs.mujoco.robots.locomotors.ant"].Ant = type("Ant", (_Any,), {})

    gym = _mod("gym")
    gym.registered = []
    envs = _mod("gym.envs")
    envs.register = lambda **kw: gym.registered.append(kw)
    gym.envs = envs
    spaces = _mod("gym.spaces")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    spaces.Box = Box
    gym.spaces = spaces
    utils = _mod("gym.utils")
    seeding = _mod("gym.utils.seeding")
    seeding.np_random = legacy_np_random
    utils.seeding = seeding
    gym.utils = utils

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return gym
