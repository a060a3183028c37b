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
        full = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
        # near-equal shards (uneven by one row when world does not divide n): sizes come from (n, world) alone
        ranges = parallel.shard_ranges(n, world)
        lo, hi = ranges[rank]
        got = parallel.all_gather_rows(full[lo:hi].clone(), ranges)
        ok = bool(torch.equal(got, full))
        # even shards: one all_gather_into_tensor straight into a preallocated output
        ne = n - n % world
        re = parallel.shard_ranges(ne, world)
        out = torch.full((ne, 3), -1.0, dtype=torch.float64)
        got_e = parallel.all_gather_rows(full[re[rank][0]:re[rank][1]], re, out=out)
        ok = ok and got_e is out and bool(torch.equal(out, full[:ne]))
        # uneven shards from the work balancer
        cost = np.linspace(1.0, 5.0, n)
        r2 = parallel.balanced_shards(cost, world)
        got2 = parallel.all_gather_rows(full[r2[rank][0]:r2[rank][1]].clone(), r2)
        ok = ok and bool(torch.equal(got2, full))
        # a rank whose rows do not match its range is refused
        try:
            parallel.all_gather_rows(full[:1], ranges)
            ok = ok and (ranges[rank][1] - ranges[rank][0] == 1)
        except ValueError:
            pass
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


class _RecordingSolver:
    """stands in for ExpertSolver (which needs a GPU): records what the sharding layer hands to it"""
    def __init__(self, dimension, nk, order, knowns, weighting_method, device=None, **kwargs):
        self.dimension, self.nk, self.order, self.knowns, self.wm = dimension, nk, order, knowns, weighting_method
        self.device, self.kwargs, self.calls = device, kwargs, []

    def prepare(self, xi, xk):
        self.calls.append(("prepare", xi, xk))

    def solve(self, fk, fi, sens=None):
        self.calls.append(("solve", fk, fi, sens))
        fi[:, 0] = fk[:, 0]          # in-place semantics of the real solve: the caller's rows are updated
        return 0

    def prepare_hoods(self, x, hoods, xi=None):
        self.calls.append(("prepare_hoods", x, hoods, xi))


def test_sharded_solver_slices_global_arrays(monkeypatch):
    """ShardedExpertSolver: every rank builds its solver from its contiguous slice of the GLOBAL metadata and hands it
    views (no copies) of its rows of the global arrays, so that in-place results land in the caller's arrays; the
    slices of all ranks tile the batch"""
    monkeypatch.setattr(parallel, "ExpertSolver", _RecordingSolver)
    n, k, world = 1001, 6, 4
    rng = np.random.default_rng(0)
    nk = rng.integers(3, k + 1, n).astype(np.int32)
    od = rng.integers(0, 5, n).astype(np.int32)
    kn = rng.integers(0, 2, n).astype(np.int64)
    wm = rng.integers(1, 3, n).astype(np.int32)
    xi, xk = rng.random((n, 2)), rng.random((n, k, 2))
    fk, fi = rng.random((n, k)), np.zeros((n, 15))
    covered = np.zeros(n, dtype=int)
    for balance in (False, True):
        fi[:] = 0.0
        covered[:] = 0
        for rank in range(world):
            s = parallel.ShardedExpertSolver(2, nk, od, kn, wm, rank=rank, world=world, device=rank, balance=balance,
                                             algorithm=1, do_sens=False)
            lo, hi = s.lo, s.hi
            covered[lo:hi] += 1
            assert s.solver.device == rank and s.solver.kwargs == {"algorithm": 1, "do_sens": False}
            assert np.array_equal(s.solver.nk, nk[lo:hi]) and np.array_equal(s.solver.order, od[lo:hi])
            assert np.array_equal(s.solver.knowns, kn[lo:hi]) and np.array_equal(s.solver.wm, wm[lo:hi])
            s.prepare(xi, xk)
            s.solve(fk, fi)
            (_, pxi, pxk), (_, sfk, sfi, ssens) = s.solver.calls
            assert np.shares_memory(pxi, xi) and np.shares_memory(pxk, xk) and np.shares_memory(sfi, fi)
            assert pxi.shape[0] == hi - lo and sfk.shape[0] == hi - lo and ssens is None
            # rows that already are local pass through untouched
            s.solve(fk[lo:hi], fi[lo:hi], local=True)
            assert s.solver.calls[-1][1].shape[0] == hi - lo
            # neighbourhoods as index lists: the cloud is replicated, the lists and the origins are this rank's rows
            hoods = rng.integers(0, n, (n, k)).astype(np.int32)
            s.prepare_hoods(xi, hoods)
            _, hx, hh, hxi = s.solver.calls[-1]
            assert hx is xi and hh.shape[0] == hi - lo and np.array_equal(hxi, xi[lo:hi])
        assert (covered == 1).all()
        assert np.array_equal(fi[:, 0], fk[:, 0])       # every rank wrote its rows of the global array in place
    # the work balancer moves the cuts towards the expensive cases
    cost = parallel.case_cost(2, nk, od, kn)
    cuts = parallel.balanced_shards(cost, world)
    work = np.array([cost[lo:hi].sum() for lo, hi in cuts])
    assert work.max() / work.mean() < 1.1
