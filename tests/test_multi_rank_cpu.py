"""World-size-2 gloo test of the N>1 host logic of bench.py: independent pairs shard across ranks
without any data-path collective; the only exchanges are the barrier and the max/sum reductions of
the timing and counters."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import bench

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = bench.rank_seeds(rank, 4)
    gathered = [None] * world
    dist.all_gather_object(gathered, seeds)
    times, sums = bench.reduce_over_ranks(dist, "cpu", [0.5 + rank, 2.0 - rank, 1.0], [10.0 * (rank + 1), 1.0, 0.0])
    dist.barrier()
    if rank == 0:
        q.put((gathered, times, sums))
    dist.destroy_process_group()


def test_pairs_shard_without_overlap_and_reductions():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, times, sums = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allseeds = sum(gathered, [])
    assert len(set(allseeds)) == len(allseeds) == 8 and min(allseeds) == 1000 and max(allseeds) == 1007
    assert times == [1.5, 2.0, 1.0]     # max over ranks
    assert sums == [30.0, 2.0, 0.0]     # sum over ranks
