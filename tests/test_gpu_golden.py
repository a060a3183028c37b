"""GPU vs the reference's golden vectors (tests/golden/, produced by the unmodified reference) and the
reference's own known-answer tests, run through the drop-in Python API -> C ABI -> CUDA kernels.

Reference tests mirrored: tests/test_simple.py (polynomial recovery, many == loop), tests/test_expert.py,
tests/test_edge_cases.py, tests/test_stencil.py, tests/test_interp.py, tests/test_parallel.py.
"""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu
wlsqm = pytest.importorskip("wlsqm_b200")

ATOL_EXACT = 1e-10   # reference tests/test_simple.py:24


@pytest.mark.parametrize("name", parity.golden_names())
def test_gpu_matches_reference_golden(name):
    c = parity.golden_case(name)
    s = wlsqm.ExpertSolver(c["dim"], c["nk"], c["od"], c["kn"], c["w"], algorithm=c["algo"], do_sens=True,
                           max_iter=c["max_iter"], ntasks=1, debug=True)
    s.prepare(c["x"], c["xk"])
    fi = c["fi0"].copy()
    sens = np.zeros((c["n"], c["k"], fi.shape[1]))
    it = s.solve(c["fk"], fi, sens)
    print(parity.check_against_floor(fi, c["fi_ref"], c["fi_ref_perm"], c["dim"], c["order"], name))
    assert it == int(c["iters_ref"])
    ns = c["sens_ref"].shape[0]
    parity.check_sens(sens[:ns], c["sens_ref"], name)
    assert np.allclose(s.conds(), c["conds_ref"], rtol=1e-5)
    # interpolation: model index from the kd-tree is the reference's, every derivative slot matches
    s.prep_interpolate()
    out0, I = s.interpolate(c["xq"], mode='nearest', diff=0)
    assert np.array_equal(I, c["I_ref"])
    # evaluate the reference's coefficients with our kernel: isolates the evaluator from the fit noise
    s2 = wlsqm.ExpertSolver(c["dim"], c["nk"], c["od"], np.full(c["n"], (1 << fi.shape[1]) - 1, np.int64), c["w"])
    s2.prepare(c["x"], c["xk"])
    fi_all_known = c["fi_ref"].copy()
    s2.solve(c["fk"], fi_all_known)          # every DOF known: the solver just adopts fi_ref
    assert np.array_equal(fi_all_known, c["fi_ref"])
    s2.tree = s.tree
    for d in range(c["interp_ref"].shape[0]):
        out, _ = s2.interpolate(c["xq"], mode='nearest', diff=d, I=I)
        ref = c["interp_ref"][d]
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(out - ref).max() / scale < 1e-12, (name, d, np.abs(out - ref).max() / scale)


def test_readme_example_and_lu_golden():
    k = np.load(parity.GOLDEN_DIR / "golden_kat.npz")
    fi = np.zeros(6)
    wlsqm.fit_2D(k["readme_xk"], k["readme_fk"], k["readme_xi"], fi, None, do_sens=0, order=2, knowns=0,
                 weighting_method=wlsqm.WEIGHT_UNIFORM)
    assert np.allclose(fi, k["readme_expected"], atol=ATOL_EXACT)
    assert np.allclose(fi, k["readme_fi"], atol=1e-12)
    from wlsqm_b200.utils import lapackdrivers as ld
    for n in (3, 15, 36):
        LU = k[f"lu{n}_A"].copy(order='F')
        ipiv = np.zeros_like(k[f"lu{n}_ipiv"])
        ld.mgeneralfactor(LU, ipiv)
        assert np.array_equal(ipiv, k[f"lu{n}_ipiv"])
        assert np.allclose(LU, k[f"lu{n}_LU"], rtol=1e-9, atol=1e-12)
        x = k[f"lu{n}_b"].copy(order='F')
        ld.mgeneralfactored(LU, ipiv, x)
        assert np.allclose(x, k[f"lu{n}_x"], rtol=1e-9, atol=1e-11)


# ---- polynomial recovery (reference tests/test_simple.py:39-129, test_edge_cases.py:34-59) --------------
def _monomials(dim, order, d):
    from wlsqm_b200.fitter import defs
    from math import factorial
    cols = []
    for (a, b, c) in defs.SLOT_EXPONENTS[dim][:defs.NUMBER_OF_DOFS[dim][order]]:
        v = d[..., 0] ** a / factorial(a)
        if dim >= 2:
            v = v * d[..., 1] ** b / factorial(b)
        if dim >= 3:
            v = v * d[..., 2] ** c / factorial(c)
        cols.append(v)
    return np.stack(cols, axis=-1)


@pytest.mark.parametrize("dim,order,nk,tol", [(1, 2, 8, 1e-10), (2, 2, 20, 1e-10), (3, 2, 40, 1e-10), (2, 3, 30, 1e-10),
                                              (2, 4, 40, 1e-8), (1, 4, 12, 1e-8), (3, 3, 60, 1e-8), (3, 4, 90, 1e-6)])
@pytest.mark.parametrize("wm", [1, 2])
def test_exact_polynomial_recovery(rng, dim, order, nk, wm, tol):
    no = wlsqm.number_of_dofs(dim, order)
    coef = rng.uniform(-2, 2, no)
    xi = rng.uniform(-1, 1, dim)
    xk = xi + rng.uniform(-0.5, 0.5, (nk, dim))
    fk = _monomials(dim, order, xk - xi) @ coef
    fi = np.zeros(no)
    fit = {1: wlsqm.fit_1D, 2: wlsqm.fit_2D, 3: wlsqm.fit_3D}[dim]
    if dim == 1:
        rc = fit(np.ascontiguousarray(xk[:, 0]), fk, float(xi[0]), fi, None, do_sens=0, order=order, knowns=0,
                 weighting_method=wm)
    else:
        rc = fit(xk, fk, xi, fi, None, do_sens=0, order=order, knowns=0, weighting_method=wm)
    assert rc == 0
    assert np.allclose(fi, coef, atol=tol * max(1.0, np.abs(coef).max())), np.abs(fi - coef).max()
    # interpolate_fit reproduces the polynomial and its derivatives (reference tests/test_interp.py:17-106)
    xq = xi + rng.uniform(-0.3, 0.3, (7, dim))
    val = wlsqm.interpolate_fit(xi if dim > 1 else float(xi[0]), fi, dim, order, xq if dim > 1 else xq[:, 0], diff=0)
    assert np.allclose(val, _monomials(dim, order, xq - xi) @ coef, atol=100 * tol)
    lam = wlsqm.lambdify_fit(xi if dim > 1 else float(xi[0]), fi, dim, order)
    assert np.allclose(lam(*[xq[:, j] for j in range(dim)]), val, rtol=1e-13, atol=1e-13)


def test_known_value_is_preserved_and_used(rng):
    # reference tests/test_edge_cases.py:62-82
    xk = rng.uniform(-1, 1, (12, 2))
    fk = rng.standard_normal(12)
    fi = np.zeros(6)
    fi[0] = 999.0
    wlsqm.fit_2D(xk, fk, np.zeros(2), fi, None, do_sens=0, order=2, knowns=wlsqm.b2_F)
    assert fi[0] == 999.0
    assert np.all(np.isfinite(fi))


def test_order0_is_mean(rng):
    xk = rng.uniform(-1, 1, (9, 2))
    fk = rng.standard_normal(9)
    fi = np.zeros(1)
    wlsqm.fit_2D(xk, fk, np.zeros(2), fi, None, do_sens=0, order=0, knowns=0, weighting_method=wlsqm.WEIGHT_UNIFORM)
    assert abs(fi[0] - fk.mean()) < 1e-12


def test_classical_stencils():
    # reference tests/test_stencil.py:58-212
    h = 0.1
    f = np.sin
    fi = np.zeros(3)
    xk = np.array([-h, 0.0, h])
    wlsqm.fit_1D(xk, f(xk), 0.0, fi, None, do_sens=0, order=2, knowns=0, weighting_method=wlsqm.WEIGHT_UNIFORM)
    assert abs(fi[0] - f(0.0)) < 1e-12 and abs(fi[1] - (f(h) - f(-h)) / (2 * h)) < 1e-11
    assert abs(fi[2] - (f(h) - 2 * f(0.0) + f(-h)) / h ** 2) < 1e-10
    # 2D five-point plus, XY eliminated as a known (= 0): the four-neighbour Laplacian stencil
    g = lambda x, y: np.exp(0.5 * x) * np.cos(y)
    xk2 = np.array([[0.0, 0.0], [h, 0.0], [-h, 0.0], [0.0, h], [0.0, -h]])
    fi2 = np.zeros(6)
    sens = np.zeros((5, 6))
    wlsqm.fit_2D(xk2, g(xk2[:, 0], xk2[:, 1]), np.zeros(2), fi2, sens, do_sens=1, order=2, knowns=wlsqm.b2_XY,
                 weighting_method=wlsqm.WEIGHT_UNIFORM)
    assert abs(fi2[wlsqm.i2_X2] - (g(h, 0) - 2 * g(0, 0) + g(-h, 0)) / h ** 2) < 1e-9
    assert abs(fi2[wlsqm.i2_Y2] - (g(0, h) - 2 * g(0, 0) + g(0, -h)) / h ** 2) < 1e-9
    assert np.isnan(sens[:, wlsqm.i2_XY]).all()
    assert np.allclose(sens[:, wlsqm.i2_X2], np.array([-2, 1, 1, 0, 0]) / h ** 2, atol=1e-8)


def test_prepare_once_solve_twice_and_errors():
    # reference tests/test_expert.py:92-117 + the exception surface of expert.pyx:131-159,493-494
    n, k = 50, 12
    x, hoods, f = parity.make_case(n, 2, k, unit_box=True)
    nk = np.full(n, k, np.int32)
    od = np.full(n, 2, np.int32)
    kn = np.zeros(n, np.int64)
    w = np.full(n, 2, np.int32)
    with pytest.raises(ValueError):
        wlsqm.ExpertSolver(4, nk, od, kn, w)
    with pytest.raises(ValueError):
        wlsqm.ExpertSolver(2, nk, od[:-1], kn, w)
    with pytest.raises(ValueError):
        wlsqm.ExpertSolver(2, nk, od, kn, w, algorithm=7)
    with pytest.raises(ValueError):
        wlsqm.ExpertSolver(2, nk, od, kn, w, ntasks=0)
    with pytest.raises(ValueError):
        wlsqm.ExpertSolver(2, nk.astype(np.int64), od, kn, w)
    s = wlsqm.ExpertSolver(2, nk, od, kn, w, ntasks=3)
    with pytest.raises(RuntimeError):
        s.solve(f[hoods], np.zeros((n, 6)))
    with pytest.raises(RuntimeError):
        s.prep_interpolate()
    s.prepare(x, x[hoods])
    with pytest.raises(RuntimeError):
        s.conds()
    with pytest.raises(RuntimeError):
        s.interpolate(x)
    fi1, fi2 = np.zeros((n, 6)), np.zeros((n, 6))
    assert s.solve(f[hoods], fi1) == 0
    assert s.solve(2.0 * f[hoods], fi2) == 0
    assert np.allclose(fi2, 2.0 * fi1, rtol=1e-12, atol=1e-12)       # linearity in the data
    used, total = s.memory_used()
    assert used > 0 and total >= used
    with pytest.raises(ValueError):
        s.solve(f[hoods].astype(np.float32), fi1)
    # guest mode shares the geometry of a prepared host (expert.pyx:163-189, 348-385)
    g = wlsqm.ExpertSolver(2, nk, od, kn, w, host=s)
    g.prepare(None, None)
    fi3 = np.zeros((n, 6))
    g.solve(f[hoods], fi3)
    assert np.array_equal(fi3, fi1)
