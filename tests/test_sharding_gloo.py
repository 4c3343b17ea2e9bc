"""world_size-2 `gloo` test of the N>1 host logic of bench.py (CPU only): image shards are disjoint and complete,
per-rank seeds never collide, and the end-of-run reduction is max over ranks for times / sum for counters."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200ocr import sharding
    mine = sharding.shard_indices(101, rank, world)
    seeds = [sharding.card_seed(rank, s, i) for s in range(4) for i in range(64)]
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, seeds))
    ms, wall, launches, words = sharding.reduce_run(dist, "cpu", 10.0 + rank, 20.0 - rank, 100 * (rank + 1), 7 + rank)
    if rank == 0:
        q.put((gathered, ms, wall, launches, words))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    import b200ocr  # noqa: F401  (the package must import without a GPU)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, ms, wall, launches, words = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = [set(g[0]) for g in gathered]
    assert shards[0].isdisjoint(shards[1]) and shards[0] | shards[1] == set(range(101))
    assert abs(len(shards[0]) - len(shards[1])) <= 1
    all_seeds = gathered[0][1] + gathered[1][1]
    assert len(set(all_seeds)) == len(all_seeds)
    assert (ms, wall, launches, words) == (11.0, 20.0, 300, 15)
