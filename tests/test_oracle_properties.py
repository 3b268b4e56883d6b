"""Size-independent properties of the task-layer arithmetic, checked on the oracle with random inputs (hypothesis).
The golden vectors (tests/test_oracle_golden.py) pin values; these pin STRUCTURE the reference's formulas imply:

* Gather sensor (`ant_gather_env.py:128-177`): a rigid motion of robot + items leaves the readings alone; the order of
  the items inside their class does not matter; nothing beyond sqrt(sensor_range) is seen; a reading is
  1 - d^2 / sensor_range of the NEAREST item of its bin (quirk: squared distance against an unsquared range).
* Wall lidar (`sizeable_enclosed_scene.py:63-97`): with span 2 pi, turning by one ray spacing shifts the readings by one
  ray; readings lie in [0, 1]; the scene's quarter-turn symmetry (square arena, centred robot) shows in the readings.
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle as O

NB, RANGE, SPAN = 10, 20.0, np.pi


def _items(rng, spread=6.0):
    return rng.uniform(-spread, spread, size=(16, 2))


def _away_from_edges(xy, yaw, items, n_bins=NB, span=SPAN, rng2=RANGE, eps=1e-6):
    """False when some item sits on a decision boundary (bin edge, half-span, range), where a rounding difference of
    the transformed problem may legitimately flip a discrete outcome."""
    d = items - xy
    ang = (np.arctan2(d[:, 1], d[:, 0]) - yaw + np.pi) % (2 * np.pi) - np.pi
    pos = (ang + span / 2) / (span / n_bins)
    d2 = (d ** 2).sum(1)
    return (np.abs(pos - np.round(pos)).min() > eps and np.abs(np.abs(ang) - span / 2).min() > eps
            and np.abs(d2 - rng2).min() > eps and np.abs(np.abs(ang) - np.pi).min() > eps)


@settings(max_examples=150, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_sensor_is_invariant_under_rigid_motion(seed):
    rng = np.random.default_rng(seed)
    items, xy, yaw = _items(rng), rng.uniform(-3, 3, 2), rng.uniform(-np.pi, np.pi)
    th, t = rng.uniform(-np.pi, np.pi), rng.uniform(-5, 5, 2)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    items2, xy2, yaw2 = items @ R.T + t, R @ xy + t, yaw + th
    if not (_away_from_edges(xy, yaw, items) and _away_from_edges(xy2, yaw2, items2)):
        return
    f1, p1, b1 = O.gather_sensor(NB, RANGE, SPAN, xy, yaw, items)
    f2, p2, b2 = O.gather_sensor(NB, RANGE, SPAN, xy2, yaw2, items2)
    assert np.array_equal(b1, b2)
    assert np.allclose(f1, f2, atol=1e-9) and np.allclose(p1, p2, atol=1e-9)


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_sensor_ignores_item_order_and_far_items(seed):
    rng = np.random.default_rng(seed)
    items, xy, yaw = _items(rng), rng.uniform(-3, 3, 2), rng.uniform(-np.pi, np.pi)
    perm = np.concatenate([rng.permutation(8), 8 + rng.permutation(8)])   # food stays food, poison stays poison
    f1, p1, _ = O.gather_sensor(NB, RANGE, SPAN, xy, yaw, items)
    f2, p2, _ = O.gather_sensor(NB, RANGE, SPAN, xy, yaw, items[perm])
    assert np.array_equal(f1, f2) and np.array_equal(p1, p2)
    # pushing every item beyond sqrt(sensor_range) blanks the sensor
    far = xy + (items - xy) / np.linalg.norm(items - xy, axis=1, keepdims=True) * (np.sqrt(RANGE) + 0.01 + rng.uniform(0, 3, (16, 1)))
    f3, p3, _ = O.gather_sensor(NB, RANGE, SPAN, xy, yaw, far)
    assert not f3.any() and not p3.any()


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_sensor_reading_is_the_nearest_item_of_the_bin(seed):
    rng = np.random.default_rng(seed)
    items, xy, yaw = _items(rng, 4.0), rng.uniform(-1, 1, 2), rng.uniform(-np.pi, np.pi)
    f, p, bins = O.gather_sensor(NB, RANGE, SPAN, xy, yaw, items)
    d2 = ((items - xy) ** 2).sum(1)
    for cls, out in ((slice(0, 8), f), (slice(8, 16), p)):
        expect = np.zeros(NB)
        for b, dd in zip(bins[0][cls], d2[cls]):
            if b >= 0:
                expect[b] = max(expect[b], 1.0 - dd / RANGE)
        assert np.allclose(out[0], expect, atol=1e-12)
    assert ((f >= 0) & (f <= 1)).all() and ((p >= 0) & (p <= 1)).all()


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_lidar_shifts_by_one_ray_per_ray_spacing(seed):
    from hrl_pybullet_envs_b200 import config as K
    rng = np.random.default_rng(seed)
    cfg = O.default_config(K.HRL_ANT_MAZE, 1)
    bounds = O.scene_bounds(cfg)
    n = 10
    xy, yaw = np.array([rng.uniform(1.5, 4.5), rng.uniform(-8, 8)]), rng.uniform(-np.pi, np.pi)
    a = O.sense_walls(n, 2 * np.pi, 5.0, bounds, xy, yaw)[0]
    b = O.sense_walls(n, 2 * np.pi, 5.0, bounds, xy, yaw + 2 * np.pi / n)[0]
    assert ((a >= 0) & (a <= 1)).all()
    # ray i of the turned robot is ray i + 1 of the original one (ray angle = pi/2 + yaw + (i + 1) / n * 2 pi)
    assert np.allclose(b, np.roll(a, -1), atol=1e-9)


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_lidar_quarter_turn_symmetry_of_the_square_arena(seed):
    from hrl_pybullet_envs_b200 import config as K
    rng = np.random.default_rng(seed)
    cfg = O.default_config(K.HRL_ANT_FLAGRUN, 1)   # 12 x 12 arena, walls only
    bounds = O.scene_bounds(cfg)
    assert len(bounds) == 4
    n = 12
    xy, yaw = rng.uniform(-4, 4, 2), rng.uniform(-np.pi, np.pi)
    xy_q = np.array([-xy[1], xy[0]])               # the same place after turning the world by 90 degrees
    a = O.sense_walls(n, 2 * np.pi, 5.0, bounds, xy, yaw)[0]
    b = O.sense_walls(n, 2 * np.pi, 5.0, bounds, xy_q, yaw + np.pi / 2)[0]
    assert np.allclose(a, b, atol=1e-9)
