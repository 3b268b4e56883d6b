"""N>1 host logic on CPU: two gloo ranks, each owning one shard of envs (stand-in stepper = the
oracle, since there is no GPU here).  Checks the shard rule (env_index_offset => the sharded job
equals the single-process batch bit-for-bit) and the two collectives bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hrl_pybullet_envs_b200.sharding import max_over_ranks, shard_offset, sum_episode_stats


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_per, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    cfg = O.default_config(0, n_per)
    cfg.seed = 7
    cfg.env_index_offset = shard_offset(rank, n_per)
    env = O.OracleVecEnv(cfg)
    obs = [env.reset()]
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (steps, n_per * world, 8)).astype(np.float32)
    for t in range(steps):
        o, r, d, info = env.step(acts[t, rank * n_per:(rank + 1) * n_per])
        obs.append(o)
    tmax = max_over_ranks([1.0 + rank, 5.0 - rank])
    f, i = env.get_state()
    tot = sum_episode_stats(i[:, 1].sum(), i[:, 2].sum())
    dist.barrier()
    q.put((rank, np.stack(obs), tmax, tot))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_shards_reproduce_single_batch():
    world, n_per, steps = 2, 6, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_per, steps, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=150) for _ in ps], key=lambda x: x[0])
    for p in ps:
        p.join(timeout=30)
        assert p.exitcode == 0
    from oracle import oracle as O
    cfg = O.default_config(0, n_per * world); cfg.seed = 7
    env = O.OracleVecEnv(cfg)
    obs = [env.reset()]
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (steps, n_per * world, 8)).astype(np.float32)
    for t in range(steps):
        obs.append(env.step(acts[t])[0])
    full = np.stack(obs)
    sharded = np.concatenate([res[0][1], res[1][1]], axis=1)
    assert np.array_equal(full, sharded)
    for r in res:
        assert r[2] == [2.0, 5.0]                      # MAX over ranks
        assert r[3] == [float(n_per * world), float(n_per * world * steps)]  # SUM over ranks
