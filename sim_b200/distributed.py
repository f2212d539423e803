"""Multi-GPU plumbing: replications shard across ranks (one process per GPU, contiguous blocks of
global replication indices; the RNG key is the GLOBAL index, so results do not depend on the number
of GPUs), and the only collective of the path is the all-reduce of the output statistics
(SURVEY.md 8(e)): ``[n, sum of means]`` and then, to keep the reference's two-pass population
variance (sim/src/output_analysis/mod.rs:23-40), ``[sum of squared deviations]``.

torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is plumbing only."""
from __future__ import annotations

from typing import Optional, Tuple


def shard_replications(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(count, offset) of the contiguous block of global replication indices owned by ``rank``."""
    base, extra = divmod(int(total), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return count, offset


def allreduce_independent_sample(shard, alpha: float, group=None, device=None) -> dict:
    """``IndependentSample::post(all replication means).confidence_interval_mean(alpha)``
    (output_analysis/mod.rs:94-125) over the shards of every rank.

    ``shard`` provides ``stat_partial() -> (sum_of_means, n)`` and ``stat_partial_sqdev(mean) -> float``
    (a ``sim_b200.Simulation``, or any stand-in with the same two methods)."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from .simulator import _CI, _check, load_library

    s, n = shard.stat_partial()
    t = torch.tensor([float(s), float(n)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    total_n = int(round(float(t[1].item())))
    if total_n == 0:
        raise ValueError("no replication has a defined mean")
    mean = float(t[0].item()) / total_n
    q = torch.tensor([float(shard.stat_partial_sqdev(mean))], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(q, op=dist.ReduceOp.SUM, group=group)
    variance = float(q.item()) / total_n
    ci = _CI()
    _check(load_library().simb_oa_confidence_interval(mean, variance, total_n, float(alpha), C.byref(ci)))
    return {"mean": ci.mean, "variance": ci.variance, "lower": ci.lower, "upper": ci.upper, "n": ci.n}
