"""CPU, world size 2 over gloo: the N>1 host logic -- index sharding, per-rank seeds and the
whole-job rate (all ranks' QPs / max-over-ranks time).  No data-path collective exists."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mgqp_b200 import sharding, workloads


def test_shard_ranges_partition_the_batch():
    for batch in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(batch, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(10_001, rank, world)
        b = workloads.make("C2", 64, seed=sharding.rank_seed(1236, rank))
        # rank 1 pretends to be slower: the job rate must be governed by the slowest rank
        rate = sharding.whole_job_rate(hi - lo, 1.0 + rank, None)
        chk = torch.tensor([float(b.g0.sum())], dtype=torch.float64)
        gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, chk)
        out.put((rank, lo, hi, rate, [float(g.item()) for g in gathered]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, rate0, g0), (r1, lo1, hi1, rate1, g1) = res
    assert (lo0, hi1) == (0, 10_001) and hi0 == lo1
    assert rate0 == rate1 == pytest.approx(10_001 / 2.0)     # max time = 2 s
    assert g0 == g1 and g0[0] != g0[1]                        # different seeds -> different shards
