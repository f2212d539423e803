"""N > 1 host logic on CPU: two gloo ranks shard replications, all-reduce the statistics partials,
and must reproduce the single-process IndependentSample of the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Shard:
    """Stand-in for a Simulation shard: the same two partial-sum methods, computed on the host."""

    def __init__(self, means):
        self.means = np.asarray(means, np.float64)

    def stat_partial(self):
        return float(self.means.sum()), int(self.means.size)

    def stat_partial_sqdev(self, mean):
        return float(((self.means - mean) ** 2).sum())


def _worker(rank, world, port, means, alpha, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sim_b200.distributed import allreduce_independent_sample, shard_replications
    count, offset = shard_replications(len(means), world, rank)
    ci = allreduce_independent_sample(_Shard(means[offset:offset + count]), alpha)
    if rank == 0:
        out.put((ci, count, offset))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_replications_partitions_exactly():
    from sim_b200.distributed import shard_replications
    for total in (1, 7, 64, 1 << 20, (1 << 26) + 3):
        for world in (1, 2, 4, 8):
            blocks = [shard_replications(total, world, r) for r in range(world)]
            assert sum(c for c, _ in blocks) == total
            assert blocks[0][1] == 0
            for (c0, o0), (c1, o1) in zip(blocks, blocks[1:]):
                assert o1 == o0 + c0


def test_two_rank_allreduce_matches_single_process():
    import oracle_api as O
    rng = np.random.default_rng(3)
    means = rng.exponential(30.0, size=1001)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, means, 0.05, out)) for r in range(2)]
    for p in procs:
        p.start()
    ci, count, offset = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    e, want = O.independent_sample(means, 0.05)
    assert (count, offset) == (501, 0) and ci["n"] == 1001
    for k in ("mean", "variance", "lower", "upper"):
        assert abs(ci[k] - want[k]) <= 1e-12 * max(1.0, abs(want[k])), k
