"""Multi-GPU plumbing: contiguous case ranges per rank, no data-path collective.

Every case of a batch is independent and all per-case inputs are row-sliced, so one
process per GPU owns an ``ExpertSolver`` over its contiguous slice ``[lo, hi)`` of the cases and
``prepare`` / ``solve`` / ``interpolate`` need no exchange at all.  The only collective is the optional
all-gather of the result rows for callers that want the whole ``fi`` on every rank
(``torch.distributed``: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from .fitter.expert import ExpertSolver
from .fitter import defs

__all__ = ["shard_range", "balanced_shards", "all_gather_rows", "ShardedExpertSolver"]


def shard_range(n: int, rank: int, world: int):
    """contiguous, near-equal split of n cases: rank r owns [lo, hi)"""
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_shards(cost, world: int):
    """contiguous split of heterogeneous cases balancing sum(cost) per rank (cost ~ nr^2 nk per case)
    -> list of (lo, hi).  SURVEY.md 8e: balance shards by work, not by case count."""
    cost = np.asarray(cost, dtype=np.float64)
    n = len(cost)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        cuts.append(int(min(max(np.searchsorted(cum, target), cuts[-1]), n)))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def case_cost(dimension, nk, order, knowns):
    """work estimate per case for balanced_shards: nr^2 * nk"""
    no = np.array([defs.NUMBER_OF_DOFS[dimension][int(o)] for o in np.asarray(order)])
    nkn = np.array([bin(int(k) & ((1 << int(m)) - 1)).count("1") for k, m in zip(np.asarray(knowns), no)])
    nr = no - nkn
    return (nr.astype(np.float64) ** 2) * np.asarray(nk, dtype=np.float64)


def all_gather_rows(local, n_total: int, lo: int, group=None):
    """Gather row slices of every rank into the full (n_total, ...) tensor on every rank.
    `local` is this rank's rows [lo, lo + len(local)); shards may be uneven (padded to the longest)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    counts[dist.get_rank(group)] = local.shape[0]
    dist.all_reduce(counts, group=group)
    longest = int(counts.max().item())
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    buf = torch.empty((world * longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    start = 0
    for r in range(world):
        c = int(counts[r].item())
        out[start:start + c] = buf[r * longest:r * longest + c]
        start += c
    assert start == n_total, "shards do not cover the batch"
    return out


class ShardedExpertSolver:
    """An ``ExpertSolver`` over this rank's contiguous slice of a global batch.

    Takes the GLOBAL metadata arrays; ``prepare`` / ``solve`` / ``interpolate`` take either global arrays
    (sliced here, no copy) or arrays that already hold just the local rows (``local=True``)."""

    def __init__(self, dimension, nk, order, knowns, weighting_method, *, rank, world, device=None, balance=False,
                 **kwargs):
        nk, order = np.asarray(nk), np.asarray(order)
        knowns, weighting_method = np.asarray(knowns), np.asarray(weighting_method)
        self.n_total = len(nk)
        if balance:
            self.lo, self.hi = balanced_shards(case_cost(dimension, nk, order, knowns), world)[rank]
        else:
            self.lo, self.hi = shard_range(self.n_total, rank, world)
        sl = slice(self.lo, self.hi)
        self.rank, self.world = rank, world
        self.solver = ExpertSolver(dimension, np.ascontiguousarray(nk[sl]), np.ascontiguousarray(order[sl]),
                                   np.ascontiguousarray(knowns[sl]), np.ascontiguousarray(weighting_method[sl]),
                                   device=device, **kwargs)

    def _rows(self, a, local):
        return a if (local or a is None) else a[self.lo:self.hi]

    def prepare(self, xi, xk, local=False):
        return self.solver.prepare(self._rows(xi, local), self._rows(xk, local))

    def solve(self, fk, fi, sens=None, local=False):
        return self.solver.solve(self._rows(fk, local), self._rows(fi, local), self._rows(sens, local))

    # -- neighbourhoods as index lists: replicated points, one all-gather of the per-point data per step --
    def prepare_hoods(self, x, hoods, local=False):
        """``x``: the FULL point cloud, replicated on every rank (16 MB per million 2D points; the neighbours of a
        contiguous case range may lie anywhere in index space unless the points are spatially sorted -- the "halo");
        ``hoods``: the global (n_total, k) int32 index lists into ``x`` (or this rank's rows with ``local=True``).
        The origins of this rank's cases are ``x[lo:hi]``."""
        return self.solver.prepare_hoods(x, self._rows(hoods, local), xi=x[self.lo:self.hi])

    def solve_hoods(self, f, fi, sens=None, local=False, f_is_local=False, group=None):
        """``f``: one value per point of the full cloud (replicated), or -- ``f_is_local=True`` -- only this rank's
        slice ``f[lo:hi]`` (what a time-stepping code owns), which is all-gathered first: the one exchange of a step,
        8 bytes per point over NCCL.  ``fi`` / ``sens``: global arrays (sliced here) or local rows (``local=True``)."""
        if f_is_local:
            f = all_gather_rows(f, self.n_total, self.lo, group)
        return self.solver.solve_hoods(f, self._rows(fi, local), self._rows(sens, local))

    def gather(self, local_rows, group=None):
        """all-gather a per-rank result (e.g. the local fi tensor) into the global array"""
        return all_gather_rows(local_rows, self.n_total, self.lo, group)
