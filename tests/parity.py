"""Shared helpers for the parity tests: seeded workloads (workloads.py), the oracle, the noise-floor criterion.

Criterion: DOF-scaled error e[i,j] = |a - b| / max_i |b[:,j]|, grouped by total derivative
order d; each of (p50, p99, max) must be <= FLOOR_FACTOR x the oracle's own self-difference under a
permutation of the neighbour order (same mathematics, different summation order -- measured in the same
run), with an absolute floor of a few ulp, and additionally p99 <= 1e-10 for d <= 1 wherever the
oracle's own floor allows that (floor p99 <= 2.5e-11; 3D order 4 with unknown F does not).
"""
from __future__ import annotations

import numpy as np

import workloads as wl
import oracle as orc

FLOOR_FACTOR = 4.0
ABS_FLOOR = 2e-14


def make_case(n, dim, k, seed=42, unit_box=False):
    x = wl.cloud(n, dim, seed=seed, unit_box=unit_box)
    hoods = wl.hoods_knn(x, k)
    f = wl.field(x)
    return x, hoods, f


def gathered(x, f, hoods):
    xk = np.ascontiguousarray(x[hoods])
    fk = np.ascontiguousarray(f[hoods])
    return xk, fk


def hetero_case(dim, n=3000):
    """a batch with per-case nk / order / knowns / weighting (expert.pyx:92-104), seeded: the inputs of
    test_gpu_parity.test_heterogeneous_batch and of the golden vectors tests/golden/golden_hetero.npz"""
    import wlsqm_b200 as wlsqm
    kmax = 30 if dim == 2 else 60
    nomax = wlsqm.number_of_dofs(dim, 4)
    b_xy = wlsqm.b2_XY if dim == 2 else wlsqm.b3_XY
    x, hoods, f = make_case(n, dim, kmax)
    rng = np.random.default_rng(3)
    od = rng.integers(0, 5, n).astype(np.int32)
    nk = np.array([rng.integers(min(kmax, (3 * wlsqm.number_of_dofs(dim, int(o))) // 2 + 2), kmax + 1) for o in od], np.int32)
    kn = np.where(rng.random(n) < 0.5, 1, 0).astype(np.int64)
    kn[od >= 2] |= np.where(rng.random((od >= 2).sum()) < 0.3, b_xy, 0)
    wm = rng.integers(1, 3, n).astype(np.int32)
    xk, fk = gathered(x, f, hoods)
    fi0 = rng.standard_normal((n, nomax))
    fi0[:, 0] = f
    return dict(dim=dim, n=n, kmax=kmax, x=x, xk=xk, fk=fk, nk=nk, od=od, kn=kn, wm=wm, fi0=fi0)


def golden_hetero(dim):
    """the unmodified reference's fi (and the NaN pattern / a checksum of sens) for hetero_case(dim), or None"""
    path = GOLDEN_DIR / "golden_hetero.npz"
    if not path.exists():
        return None
    z = np.load(path)
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("d%d/" % dim)}


def oracle_solve(dim, nk, order, knowns, wm, xi, xk, fk, fi0, algorithm=1, do_sens=False, max_iter=10):
    s = orc.OracleSolver(dim, nk, order, knowns, wm, algorithm, do_sens, max_iter)
    s.prepare(xi, xk)
    fi = fi0.copy()
    sens = np.zeros((len(nk), xk.shape[1], fi.shape[1])) if do_sens else None
    it = s.solve(fk, fi, sens)
    return fi, sens, it, s


def permuted_self_noise(dim, nk, order, knowns, wm, xi, xk, fk, fi0, algorithm=1, max_iter=10, seed=7):
    """oracle(fk, xk) vs oracle(neighbours permuted): the reference algorithm's own reproducibility floor"""
    rng = np.random.default_rng(seed)
    k = xk.shape[1]
    perm = rng.permutation(k)
    a, _, _, _ = oracle_solve(dim, nk, order, knowns, wm, xi, xk, fk, fi0, algorithm, False, max_iter)
    b, _, _, _ = oracle_solve(dim, nk, order, knowns, wm, xi, np.ascontiguousarray(xk[:, perm]),
                              np.ascontiguousarray(fk[:, perm]), fi0, algorithm, False, max_iter)
    return a, b


def permute_within_hoods(nk, xk, fk, seed=7):
    """per-case random permutation of the first nk[j] neighbours (the rest of the row stays where it is): the same
    neighbourhoods in another summation order, for batches whose cases use different numbers of neighbours"""
    rng = np.random.default_rng(seed)
    n, kmax = fk.shape
    keys = rng.random((n, kmax))
    keys[np.arange(kmax)[None, :] >= np.asarray(nk)[:, None]] = 2.0 + np.arange(kmax)[None, :].repeat(n, 0)[
        np.arange(kmax)[None, :] >= np.asarray(nk)[:, None]]
    perm = np.argsort(keys, axis=1, kind="stable")
    rows = np.arange(n)[:, None]
    return np.ascontiguousarray(xk[rows, perm]), np.ascontiguousarray(fk[rows, perm])


def hetero_self_noise(dim, nk, order, knowns, wm, xi, xk, fk, fi0, algorithm=1, max_iter=10, seed=7, seeds=None,
                      iter_rounds=False):
    """oracle vs oracle with every case's own neighbours permuted (heterogeneous batches).  With `seeds`, the
    permuted result that lies FARTHEST from the unpermuted one, entry by entry, over several permutations: the maximum
    of a few hundred heavy-tailed errors (one worst-conditioned case decides it) is a noisy statistic of one sample.

    iter_rounds (ALGO_ITERATIVE): the reference leaves its refinement loop when the residual norm repeats BIT FOR BIT
    (impl.pyx:1057-1060), so another summation order can leave a round earlier or later; for a case whose refinement
    oscillates instead of converging (a handful of neighbours, one unknown) consecutive rounds differ by far more than
    the rounding noise.  The oracle's own results one round before and one round after max_iter belong to its
    reproducibility floor for that reason and are folded in."""
    a, _, _, _ = oracle_solve(dim, nk, order, knowns, wm, xi, xk, fk, fi0, algorithm, False, max_iter)
    far = None

    def fold(b):
        nonlocal far
        far = b if far is None else np.where(np.abs(b - a) > np.abs(far - a), b, far)
    for sd in (seeds or (seed,)):
        xkp, fkp = permute_within_hoods(nk, xk, fk, sd)
        fold(oracle_solve(dim, nk, order, knowns, wm, xi, xkp, fkp, fi0, algorithm, False, max_iter)[0])
    if iter_rounds and algorithm == 2:
        for mi in (max_iter - 1, max_iter + 1):
            if mi >= 1:
                fold(oracle_solve(dim, nk, order, knowns, wm, xi, xk, fk, fi0, algorithm, False, mi)[0])
    return a, far


def check_hetero_against_floor(got, ref, ref_perm, dim, od, label="", min_cases=20):
    """check_against_floor per ORDER GROUP of a batch with per-case orders: the cases of order o, their first no(o)
    columns, grouped by derivative order.  Returns the report."""
    lines = []
    for order in sorted(set(int(o) for o in od)):
        m = np.asarray(od) == order
        if m.sum() < min_cases:
            continue
        no = orc.number_of_dofs(dim, order)
        lines.append(check_against_floor(got[m][:, :no], ref[m][:, :no], ref_perm[m][:, :no], dim, order,
                                         "%s order %d (%d cases)" % (label, order, int(m.sum()))))
    return "\n".join(lines)


def golden_hetero_iter(dim):
    """the unmodified reference's ALGO_ITERATIVE(3) results for hetero_case(dim) (tests/golden/golden_hetero_iter.npz)"""
    path = GOLDEN_DIR / "golden_hetero_iter.npz"
    if not path.exists():
        return None
    z = np.load(path)
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("d%d/" % dim)}


def golden_continuous(dim):
    """the unmodified reference's ExpertSolver.interpolate outputs, both modes (tests/golden/golden_continuous.npz)"""
    path = GOLDEN_DIR / "golden_continuous.npz"
    if not path.exists():
        return None
    z = np.load(path)
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("d%d/" % dim)}


def check_against_floor(got, ref, ref_perm, dim, order, label=""):
    """assert got ~ ref within FLOOR_FACTOR x |ref - ref_perm| per derivative order; returns the report"""
    rep = wl.parity_report(got, ref, dim, order)
    floor = wl.parity_report(ref_perm, ref, dim, order)
    lines = []
    ok = True
    for d, (p50, p99, mx) in rep.items():
        f50, f99, fmx = floor.get(d, (0.0, 0.0, 0.0))
        lim = (FLOOR_FACTOR * f50 + ABS_FLOOR, FLOOR_FACTOR * f99 + 10 * ABS_FLOOR, FLOOR_FACTOR * fmx + 100 * ABS_FLOOR)
        good = p50 <= lim[0] and p99 <= lim[1] and mx <= lim[2]
        if d <= 1 and f99 <= 2.5e-11:     # the flat 1e-10 bar, wherever the reference itself can meet it
            good = good and p99 <= 1e-10
        ok = ok and good
        lines.append(f"{label} d{d}: got p50/p99/max = {p50:.2e}/{p99:.2e}/{mx:.2e}  floor = {f50:.2e}/{f99:.2e}/{fmx:.2e}"
                     f"  {'ok' if good else 'FAIL'}")
    report = "\n".join(lines)
    assert ok, "parity outside the reference's own noise floor:\n" + report
    return report


# ---- golden vectors (tests/golden/, generated from the unmodified reference by make_golden.py) ----------
from pathlib import Path as _Path

GOLDEN_DIR = _Path(__file__).resolve().parent / "golden"
_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = np.load(GOLDEN_DIR / "golden_cases.npz")
    return _golden


def golden_names():
    return [str(s) for s in golden()["names"]]


def golden_case(name):
    g = golden()
    c = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}
    dim, order, k, knowns, wm, algo, max_iter, n = (int(v) for v in c["meta"])
    c.update(dim=dim, order=order, k=k, knowns=knowns, wm=wm, algo=algo, max_iter=max_iter, n=n)
    c["nk"] = np.full(n, k, np.int32)
    c["od"] = np.full(n, order, np.int32)
    c["kn"] = np.full(n, knowns, np.int64)
    c["w"] = np.full(n, wm, np.int32)
    c["xk"] = np.ascontiguousarray(c["x"][c["hoods"]])
    c["fk"] = np.ascontiguousarray(c["f"][c["hoods"]])
    return c


def check_sens(sens_got, sens_ref, label=""):
    """sens parity: identical NaN pattern; entries within cond*eps of the column's largest entry"""
    assert np.array_equal(np.isnan(sens_got), np.isnan(sens_ref)), label + ": NaN pattern of sens differs"
    r = np.nan_to_num(sens_ref)
    sc = np.abs(r).max(axis=(0, 1))
    sc[sc == 0] = 1.0
    err = np.abs(np.nan_to_num(sens_got) - r) / sc
    assert err.max() < 1e-7 and np.median(err) < 1e-11, (label, err.max(), np.median(err))
    return err.max()


def cfg4_problem(dim, n, k, order=3):
    """BASELINE.json configs[3] as SURVEY.md 8d builds it (needs a GPU: the neighbour search runs on the device):
    order-3 fits with WEIGHT_UNIFORM; interior points know F; boundary points (2D: the 4 sqrt(n) points nearest the box
    edges; 1D: the two end points and every 1000th point) know dF/dy (dF/dx in 1D) from the analytic field instead and
    take their k neighbours from the INTERIOR points only (defs.pyx:199-207); an interior point's neighbours are its k
    nearest interior points.  Returns host metadata and device tensors (x_d, hoods_d, f_d, fi_in_d)."""
    import torch
    import wlsqm_b200 as wlsqm
    no = orc.number_of_dofs(dim, order)
    x = wl.cloud(n, dim)
    x2 = x.reshape(n, -1)
    L = wl.H0 * n ** (1.0 / dim)
    bnd = np.zeros(n, bool)
    if dim == 2:
        edge = np.minimum(np.minimum(x2[:, 0], L - x2[:, 0]), np.minimum(x2[:, 1], L - x2[:, 1]))
        bnd[np.argsort(edge, kind="stable")[: int(4 * np.sqrt(n))]] = True
        iB, bB, bF = wlsqm.i2_Y, wlsqm.b2_Y, wlsqm.b2_F
        dfd = -np.pi * np.sin(np.pi * x2[:, 0]) * np.sin(np.pi * x2[:, 1])       # d/dy of sin(pi x) cos(pi y)
    elif dim == 1:
        bnd[[int(np.argmin(x)), int(np.argmax(x))]] = True
        bnd[::1000] = True
        iB, bB, bF = wlsqm.i1_X, wlsqm.b1_X, wlsqm.b1_F
        dfd = np.pi * np.cos(np.pi * x)
    else:
        raise ValueError("configs[3] is 1D / 2D")
    interior = np.nonzero(~bnd)[0]
    boundary = np.nonzero(bnd)[0]
    xin_d = torch.from_numpy(np.ascontiguousarray(x2[interior])).cuda()
    grid = wlsqm.PointGrid(xin_d)
    h_in = grid.knn(k)                                                            # (n_int, k) into the interior array
    _, h_b = grid.query(torch.from_numpy(np.ascontiguousarray(x2[boundary])).cuda(), k=k)
    int_d = torch.from_numpy(interior).cuda()
    hoods_d = torch.empty((n, k), dtype=torch.int32, device="cuda")
    hoods_d[int_d] = int_d[h_in.long()].int()
    hoods_d[torch.from_numpy(boundary).cuda()] = int_d[h_b.reshape(len(boundary), k)].int()
    grid.close()
    nk, od, wm = np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, 1, np.int32)     # WEIGHT_UNIFORM
    kn = np.full(n, bF, np.int64)
    kn[bnd] = bB
    f = wl.field(x)
    fi_in = np.full((n, no), 123.0)                    # sentinel in every unknown slot
    fi_in[~bnd, 0] = f[~bnd]
    fi_in[bnd, iB] = dfd[bnd]
    return dict(dim=dim, n=n, k=k, order=order, no=no, x=x, x2=x2, bnd=bnd, interior=interior, boundary=boundary, iB=iB,
                meta=(nk, od, kn, wm), f=f, hoods_d=hoods_d, x_d=torch.from_numpy(x).cuda(), f_d=torch.from_numpy(f).cuda(),
                fi_in_d=torch.from_numpy(fi_in).cuda())
