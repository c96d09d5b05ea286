"""Utterance sharding across ranks / GPUs.

The reference deals contiguous chunks of ``ceil(N / procs)`` files to ``num_gpus * workers_per_gpu`` processes with
``gpu_id = rank % num_gpus`` and no collective (``preprocess/process_dataset.py:256-278``).  Feature extraction is
embarrassingly parallel per utterance, so the same holds here: every rank takes a disjoint set of utterances and no
data-path collective exists; only the statistics pass ends with one all-reduce (``stats.MelStatsAccumulator``).
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np


def contiguous_shard(n_items: int, rank: int, world_size: int) -> range:
    """The reference's partition: contiguous chunks of ``ceil(N / world)`` (process_dataset.py:256-259)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    chunk = math.ceil(n_items / world_size) if n_items else 0
    lo = min(n_items, rank * chunk)
    return range(lo, min(n_items, lo + chunk))


def balanced_shards(lengths: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Length-balanced assignment for variable-length clips: longest-processing-time-first greedy on the sample
    counts.  Returns, per rank, the sorted indices of its utterances; every index appears exactly once."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    loads = np.zeros(world_size, dtype=np.int64)
    buckets: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(loads))
        buckets[r].append(int(i))
        loads[r] += int(lengths[i])
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]


def batches_by_budget(lengths: Sequence[int], indices: Sequence[int], max_samples: int) -> List[np.ndarray]:
    """Group a rank's utterances into launches of at most ``max_samples`` samples (at least one clip each)."""
    out: List[np.ndarray] = []
    cur: List[int] = []
    tot = 0
    for i in indices:
        n = int(lengths[i])
        if cur and tot + n > max_samples:
            out.append(np.array(cur, dtype=np.int64))
            cur, tot = [], 0
        cur.append(int(i))
        tot += n
    if cur:
        out.append(np.array(cur, dtype=np.int64))
    return out
