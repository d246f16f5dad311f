"""not-gpu: the multi-GPU path is job dealing + one max-reduction; exercised with two gloo processes on the CPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resynthesizer_b200 import sharding  # noqa: E402


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.deal_round_robin(64, world, rank)                 # BASELINE config 5: 64 equal jobs
    t = [10.0 + rank, 3.0 - rank, float(len(mine))]
    mx = sharding.max_over_ranks(t)
    sm = sharding.sum_over_ranks([float(len(mine)), float(sum(mine))])
    dist.barrier()
    out.put((rank, mine, mx, sm))
    dist.destroy_process_group()


def test_two_ranks_deal_all_jobs_once_and_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    jobs = sorted(j for _r, mine, _m, _s in res for j in mine)
    assert jobs == list(range(64))
    for _r, mine, mx, sm in res:
        assert mx == [11.0, 3.0, 32.0]
        assert sm == [64.0, float(sum(range(64)))]


def test_lpt_balances_unequal_jobs():
    costs = [sharding.job_cost(n, 30, 200) for n in (4096, 65536, 65536, 1 << 20, 4096, 262144, 262144, 16384)]
    shares = sharding.deal_lpt(costs, 4)
    assert sorted(i for s in shares for i in s) == list(range(len(costs)))
    loads = [sum(costs[i] for i in s) for s in shares]
    assert max(loads) == costs[3]            # the biggest job sits alone
    assert sharding.deal_lpt([1.0] * 8, 8) == [[i] for i in range(8)]


def test_single_process_is_identity():
    assert sharding.max_over_ranks([1.5, 2.5]) == [1.5, 2.5]
    assert sharding.env_rank()[2] >= 1
