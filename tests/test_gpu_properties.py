"""The size-independent properties of tests/test_oracle_properties.py on the CUDA parity entry points, at sizes no CPU
oracle finishes in seconds (2^20 poses per call = 256 AntGather batches of 4096 envs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
M = 1 << 20


def _poses(seed, spread=7.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    xy = (torch.rand(M, 2, generator=g, device="cuda") * 2 - 1) * spread
    yaw = (torch.rand(M, generator=g, device="cuda") * 2 - 1) * np.pi
    return g, xy, yaw


def test_sensor_permutation_and_range_at_full_size():
    from hrl_pybullet_envs_b200.vec_env import gather_sensor
    g, xy, yaw = _poses(0)
    items = xy[:, None, :] + (torch.rand(M, 16, 2, generator=g, device="cuda") * 2 - 1) * 5
    f1, p1, b1 = gather_sensor(xy, yaw, items)
    perm = torch.cat([torch.randperm(8, generator=g, device="cuda"), 8 + torch.randperm(8, generator=g, device="cuda")])
    f2, p2, b2 = gather_sensor(xy, yaw, items[:, perm])
    # the order of the items inside their class does not matter (ant_gather_env.py:136-177: nearest wins per bin) - bit for bit
    assert torch.equal(f1, f2) and torch.equal(p1, p2) and torch.equal(b1[:, perm], b2)
    assert float(f1.min()) >= 0 and float(f1.max()) <= 1 and float(p1.min()) >= 0 and float(p1.max()) <= 1
    # a reading is 1 - d^2 / sensor_range of the nearest item of its bin (squared distance against the unsquared range)
    d2 = ((items - xy[:, None, :]).double() ** 2).sum(-1)
    val = torch.where(b1 >= 0, 1.0 - d2 / 20.0, torch.zeros_like(d2))
    for cls, out in ((slice(0, 8), f1), (slice(8, 16), p1)):
        exp = torch.zeros(M, 10, dtype=torch.float64, device="cuda")
        exp.scatter_reduce_(1, b1[:, cls].clamp(min=0).long(), val[:, cls], reduce="amax", include_self=True)
        assert float((out.double() - exp).abs().max()) <= 1e-5
    # nothing beyond sqrt(sensor_range) is seen
    d = items - xy[:, None, :]
    far = xy[:, None, :] + d / d.norm(dim=-1, keepdim=True) * (np.sqrt(20.0) + 0.01 + torch.rand(M, 16, 1, generator=g, device="cuda") * 3)
    f3, p3, b3 = gather_sensor(xy, yaw, far)
    assert not bool(f3.any()) and not bool(p3.any()) and bool((b3 < 0).all())


def _flip_rate(a, b, tol):
    return float(((a - b).abs() > tol).float().mean())


def test_lidar_symmetries_at_full_size():
    from hrl_pybullet_envs_b200.vec_env import sense_walls
    from hrl_pybullet_envs_b200 import config as K
    from oracle import oracle as O
    # (the bound lines come from the same config struct the kernels use; the oracle call only formats them)
    bounds = torch.tensor(O.scene_bounds(O.default_config(K.HRL_ANT_FLAGRUN, 1)), dtype=torch.float32)
    g, xy, yaw = _poses(1, spread=5.0)
    n = 12
    a = sense_walls(xy, yaw, bounds, n, 2 * np.pi, 5.0)
    assert float(a.min()) >= 0 and float(a.max()) <= 1
    # quarter turn of the square arena (sizeable_enclosed_scene.py:28-34): same readings from the turned pose
    xy_q = torch.stack([-xy[:, 1], xy[:, 0]], dim=1)
    b = sense_walls(xy_q, yaw + np.pi / 2, bounds, n, 2 * np.pi, 5.0)
    # float32 inputs: the turned yaw is rounded (6e-8 rad), so a reading moves by ~1e-6 and a ray grazing a decision
    # (range limit, corner between two lines) may flip; such flips must stay rare
    assert _flip_rate(a, b, 2e-5) < 2e-4, _flip_rate(a, b, 2e-5)
    # turning by one ray spacing shifts the readings by one ray (ray i: pi/2 + yaw + (i + 1) / n * 2 pi, :71-74)
    c = sense_walls(xy, yaw + 2 * np.pi / n, bounds, n, 2 * np.pi, 5.0)
    assert _flip_rate(c, torch.roll(a, -1, dims=1), 2e-5) < 2e-4
