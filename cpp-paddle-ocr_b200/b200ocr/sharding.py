"""Host-side sharding of the image stream across ranks (one process per GPU, no data-path collective).

The reference's pools hand every request to one worker (src/gpu_worker_pool.cpp:46-59); across GPUs the same holds:
images are independent, so rank r simply owns every image whose index i satisfies i % world == r.  The only
cross-rank traffic is the timing / counter reduction at the end of a benchmark run."""
from __future__ import annotations


def shard_indices(n_total: int, rank: int, world: int):
    """Indices of the images rank `rank` processes out of n_total (round-robin, like the pool's dispatch)."""
    return list(range(rank, n_total, world))


def card_seed(rank: int, set_index: int, i: int) -> int:
    """Seed of image i of input set `set_index` on `rank`: distinct across ranks, sets and images."""
    return 1_000_000 * rank + 10_000 * set_index + i


def reduce_run(dist, device, elapsed_ms: float, wall_ms: float, launches: int, words: int):
    """max over ranks of the two times, sum over ranks of the two counters (what bench.py reports)."""
    import torch
    t = torch.tensor([elapsed_ms, wall_ms], device=device, dtype=torch.float64)
    c = torch.tensor([float(launches), float(words)], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t[0]), float(t[1]), int(c[0]), int(c[1])
