"""Dealing independent synthesis jobs to GPUs.

A job never shards; a batch does, with no data-path collective.  The only communication is
one max-reduction of a few timing scalars after the work, so that reported times are the max over ranks."""
import os


def deal_round_robin(n_jobs, world, rank):
    """Indices of the jobs rank `rank` runs: equal-cost jobs dealt round robin (BASELINE config 5)."""
    return list(range(rank, n_jobs, world))


def deal_lpt(costs, world):
    """Longest-processing-time-first for unequal jobs; cost ~ n_targets * (patch + probes).
    Returns a list of `world` lists of job indices."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world
    shares = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        shares[r].append(i)
        loads[r] += costs[i]
    return shares


def job_cost(n_targets, patch, probes):
    return float(n_targets) * float(patch + probes)


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def max_over_ranks(values, device=None):
    """Element-wise max over all ranks of a list of floats (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(values, dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(values, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(values, dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]
