"""Multi-GPU plumbing: envs shard by GLOBAL index across ranks, no collective
inside `step`; ONE all-reduce (NCCL over NVLink on GPUs, gloo in CPU tests) of a
float64[8] statistics vector per match."""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total: int, rank: int, world: int):
    """env index range [lo, hi) of `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sums(sums: torch.Tensor) -> torch.Tensor:
    """The single collective of the path: sum the f64[8] partial statistics over ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def stats_from_sums(sums) -> tuple:
    """(mean, standard error with ddof=1, win rate) from {n, sum x, sum x^2, #(x>0)}
    -- src/evaluation.py:199-201 formed identically on every rank."""
    n, s1, s2, w = (float(v) for v in sums[:4])
    mean = s1 / n
    var = max(0.0, (s2 - s1 * s1 / n) / (n - 1)) if n > 1 else float("nan")
    return mean, math.sqrt(var) / math.sqrt(n), w / n
