"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding and the summary gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bpm_analysis_b200 import dist as bdist


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        bdist.shard_range(4, 2, 2)


def test_shard_by_cost_balances():
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 100, 64).astype(float)
    parts = bdist.shard_by_cost(costs, 4)
    assert sorted(i for p in parts for i in p) == list(range(64))
    loads = [costs[p].sum() for p in parts]
    assert max(loads) - min(loads) <= costs.max()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = bdist.shard_range(5, world, rank)            # 5 recordings over 2 ranks: 3 + 2
        local = np.array([[i, 10.0 * i, rank] for i in range(lo, hi)], dtype=np.float64).reshape(-1, 3)
        table = bdist.gather_summaries(local, torch.device("cpu"))
        q.put((rank, table))
    finally:
        dist.destroy_process_group()


def test_gather_summaries_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.array([[0, 0, 0], [1, 10, 0], [2, 20, 0], [3, 30, 1], [4, 40, 1]], dtype=np.float64)
    for r in range(world):
        assert np.array_equal(got[r], expect)
