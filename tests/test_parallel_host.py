"""Host-side logic of the N>1 path on CPU: shard arithmetic and the result all-gather over gloo, world size 2."""
import os
import socket

import numpy as np
import pytest

from wlsqm_b200 import parallel


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            rs = [parallel.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)


def test_balanced_shards_balance_work():
    rng = np.random.default_rng(0)
    order = rng.integers(0, 5, 5000).astype(np.int32)
    nk = rng.integers(10, 60, 5000).astype(np.int32)
    kn = rng.integers(0, 2, 5000).astype(np.int64)
    cost = parallel.case_cost(2, nk, order, kn)
    shards = parallel.balanced_shards(cost, 4)
    assert shards[0][0] == 0 and shards[-1][1] == 5000
    work = np.array([cost[lo:hi].sum() for lo, hi in shards])
    assert work.max() / work.mean() < 1.05


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = parallel.shard_range(n, rank, world)
        full = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
        got = parallel.all_gather_rows(full[lo:hi].clone(), n, lo)
        ok = bool(torch.equal(got, full))
        # uneven shards from the work balancer
        cost = np.linspace(1.0, 5.0, n)
        lo2, hi2 = parallel.balanced_shards(cost, world)[rank]
        got2 = parallel.all_gather_rows(full[lo2:hi2].clone(), n, lo2)
        ok = ok and bool(torch.equal(got2, full))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_all_gather_rows_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 101, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
