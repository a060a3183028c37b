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
