"""Minimal gym<=0.21 surface (gym itself is not installed in this image): ``Box``, ``register``,
``make`` and a single-env view over ``VecEnv`` so that the reference README loop
(README.md:20-37) runs unchanged apart from the import.

Registered ids = hrl_pybullet_envs/__init__.py:9-16 plus ``AntMjBulletEnv-v0`` (README.md:13).
If a real ``gym`` / ``gymnasium`` is importable the ids are registered there as well.
"""
import numpy as np

from .config import ENV_IDS, HRL_POINT_GATHER


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low = np.full(shape, low, dtype=dtype)
        self.high = np.full(shape, high, dtype=dtype)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


class SingleEnv:
    """N=1 view: ``reset() -> obs``, ``step(a) -> (obs, rew, done, info)`` with numpy values.
    Like gym's TimeLimit wrapper it does not auto-reset."""

    metadata = {"render.modes": []}

    def __init__(self, env_id, device=0, seed=None, **kwargs):
        from .vec_env import VecEnv
        self.vec = VecEnv(env_id, 1, device=device, seed=seed, auto_reset=False, **kwargs)
        self.observation_space = Box(-np.inf, np.inf, (self.vec.D,))
        self.action_space = Box(-1.0, 1.0, (self.vec.A,))
        self.spec = type("Spec", (), {"id": env_id, "max_episode_steps": 2000})()
        self._ctor_seed = seed   # the ctor kwarg (for Flagrun: the shared goal stream, ant_flagrun_env.py:16,39)
        self._seed = seed
        self._kwargs = kwargs

    def seed(self, seed=None):
        """gym's env.seed(s): re-keys the env's own RNG streams (joint noise, item placement, maze goal); the Flagrun
        goal stream keeps the ctor's seed, like the reference's private RandomState."""
        if seed is not None and seed != self._seed:
            from .vec_env import VecEnv
            self.vec.close()
            self.vec = VecEnv(self.spec.id, 1, device=self.vec.device.index, seed=self._ctor_seed, env_seed=seed, auto_reset=False,
                              **self._kwargs)
            self._seed = seed
        return [self._seed]

    def reset(self):
        return self.vec.reset().cpu().numpy()[0].copy()

    def step(self, a):
        # numpy in / numpy out through hrl_step_host: ONE launch, results land in pinned host memory (no torch ops,
        # no per-value device synchronisation)
        act = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(1, -1))
        obs, rew, done, info = self.vec.step_host(act)
        i = {}
        for k, v in info.items():
            if k == "TimeLimit.truncated":
                if bool(v[0]):
                    i[k] = True
            elif k == "target_switched":
                if bool(v[0]):   # ant_flagrun_env.py:188-191,199: info['target'] only in a step that switched goals
                    t = info["target"]
                    i["target"] = (float(t[0, 0]), float(t[0, 1]))
            elif k == "target":
                pass
            elif k != "terminal_obs":
                i[k] = float(v[0])
        return obs[0].copy(), float(rew[0]), bool(done[0]), i

    # AntFlagrunBulletEnv's public goal methods (ant_flagrun_env.py:57,91-120); TypeError on the other envs
    @property
    def goal(self):
        g = self.vec.goal[0]
        return (float(g[0]), float(g[1]))

    def set_target(self, x, y):
        self.vec.set_target([float(x), float(y)])

    def create_targets(self, n):
        self.vec.create_targets(n)

    def next_target(self):
        return self.vec.next_target().cpu().numpy()[0].copy()

    def render(self, *a, **k):
        return None  # headless batched simulator: no renderer

    def close(self):
        self.vec.close()


registry = {}


def register(id, entry_point=None, max_episode_steps=2000, **kw):
    registry[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps, **kw)


def make(id, **kwargs):
    if id not in registry:
        raise KeyError("No registered env with id: %s" % id)
    return SingleEnv(id, **kwargs)


for _id in ENV_IDS:
    register(_id, entry_point="hrl_pybullet_envs_b200.gym_shim:SingleEnv", max_episode_steps=2000)


def register_with_installed_gym():
    """Best effort: expose the ids through a real gym/gymnasium when one is importable."""
    done = []
    for modname in ("gym", "gymnasium"):
        try:
            mod = __import__(modname)
            for _id in ENV_IDS:
                try:
                    mod.envs.registration.register(id=_id, entry_point=lambda _id=_id, **kw: SingleEnv(_id, **kw),
                                                   max_episode_steps=2000)
                except Exception:
                    pass
            done.append(modname)
        except ImportError:
            pass
    return done


REGISTERED_WITH = register_with_installed_gym()   # [] in this image (no gym / gymnasium installed)
