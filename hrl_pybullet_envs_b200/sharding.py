"""Multi-GPU rule: envs are independent, so GPU/rank r owns the global envs
[r*n, (r+1)*n) and steps them with its own handle and stream - no collective on the step path
(SURVEY.md 8e).  The only collectives are the benchmark's max-over-ranks timing and the optional
episode-statistics sum; both work on any torch.distributed backend (NCCL on GPUs, gloo in tests).
"""
import torch


def shard_offset(rank, envs_per_rank):
    """Global index of local env 0 of `rank` (hrl_config.env_index_offset): the RNG key of an env is
    (seed, global index), so a sharded job reproduces the single-process batch bit-for-bit."""
    return int(rank) * int(envs_per_rank)


def max_over_ranks(values, device=None):
    """Element-wise MAX of a list of floats over all ranks (timings are the slowest rank's)."""
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def sum_episode_stats(episodes, steps, extra=(), device=None):
    """SUM over ranks of (episodes finished, env-steps taken, *extra): the optional statistics
    aggregation the north star allows NCCL for.  Inputs are per-rank scalars/tensors."""
    import torch.distributed as dist
    t = torch.tensor([float(episodes), float(steps)] + [float(x) for x in extra], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.cpu()]


def gather_over_ranks(values, device=None):
    """Every rank's list of floats, as a list indexed by rank (per-rank timing reports of the benchmark)."""
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return [[float(x) for x in o.cpu()] for o in out]
    return [[float(x) for x in t.cpu()]]
