"""Utterance sharding across the GPUs of one box.

Every utterance is an independent dynamic program (the reference's batch loop is an independent
``prange``, core.pyx:44-45), so the path shards by utterance with NO collective: rank r runs the
kernels on its own utterances and nothing is exchanged.  In training this is what DDP's
DistributedSampler already does (``__main__.py:235``); these helpers serve the multi-GPU
benchmark and callers that hold a whole batch on one host.
"""
from __future__ import annotations

from typing import List, Sequence


def contiguous_shard(num_utterances: int, world_size: int, rank: int) -> range:
    """Contiguous split of the batch dimension; the first ``B % world_size`` ranks get one more."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad world_size / rank")
    base, extra = divmod(num_utterances, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def balanced_shards(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment by per-utterance cost (t_x * t_y): batches come
    sorted by length (dataset.py:79-81), so a contiguous split would give rank 0 all the long ones."""
    if world_size < 1:
        raise ValueError("bad world_size")
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    for s in shards:
        s.sort()
    return shards
