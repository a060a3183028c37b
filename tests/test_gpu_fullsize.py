"""Parity at BASELINE.json's FULL sizes, through size-independent properties (the oracle finishes a 1M-point
ExpertSolver in minutes, not seconds, so at these sizes it checks a strided subsample and the rest is covered by
properties of the mathematics):

  * polynomial reproduction: data sampled from a polynomial of the model's degree is fitted exactly, at EVERY point
    (the reference's own known-answer test, tests/test_simple.py:39-129, at 1M-4M cases instead of one),
  * linearity of solve in the data,
  * shard invariance: a contiguous sub-range of cases solved on its own is bit-identical to the same rows of the
    whole batch (what the multi-GPU sharding of DESIGN.md section 5 relies on),
  * hood indexing: device-side kNN == cKDTree on a sample, gather-on-device path == pre-gathered path, bit for bit,
  * knowns left untouched bit for bit at every point; sens^T fk == fi; NaN pattern of sens,
  * interpolate: one all-slots pass == 15 reference-style calls to a few ulp; evaluation at the model origin returns
    the stored coefficient exactly,
  * the oracle on every ~1000th case, with the noise-floor criterion of tests/parity.py.

Tolerances (FP64).  Polynomial reproduction: the polynomial is scaled so that every term is O(1) over a
neighbourhood, hence every DOF is recovered to cond(A_scaled) * eps relative to its own exact value: p99 <= 1e-11
over all points and <= 1e-8 for the worst-conditioned of the 1M-4M neighbourhoods (scaled condition numbers reach
4e6-9e6; the reference's own known-answer tolerance for single well-conditioned cases is 1e-10,
tests/test_simple.py:24).  Linearity and the sens identity are measured on an oscillating field, DOF-scaled by
max_i |fi[:, j]|, against 100x the p99 of the reference's own neighbour-permutation noise by derivative order
(SURVEY.md 8c table): d0 2e-12, d1 5e-11, d2 5e-9, d3 2e-7, d4 1e-5, times 100 for the maximum over all points.
"""
import numpy as np
import pytest

import parity
import workloads as wl

pytestmark = pytest.mark.gpu

wlsqm = pytest.importorskip("wlsqm_b200")
torch = pytest.importorskip("torch")

P99_TOL = {0: 2e-12, 1: 5e-11, 2: 5e-9, 3: 2e-7, 4: 1e-5}
REPRO_P99, REPRO_MAX = 1e-11, 1e-8


def _slot_exps(dim, order):
    """exponent tuple of every DOF slot, decoded from the package's own constants i{dim}_X2Y ... (the slot order of
    wlsqm/fitter/defs.pyx:201-287 is not lexicographic in 3D)"""
    import re
    no = wlsqm.number_of_dofs(dim, order)
    out = [None] * no
    for name in dir(wlsqm):
        m = re.fullmatch(r"i%d_((?:[XYZ][2-4]?)+|F)" % dim, name)
        if not m:
            continue
        slot = getattr(wlsqm, name)
        if slot >= no:
            continue
        e = [0, 0, 0]
        if m.group(1) != "F":
            for ax, p in re.findall(r"([XYZ])([2-4]?)", m.group(1)):
                e["XYZ".index(ax)] += int(p) if p else 1
        assert all(v == 0 for v in e[dim:]) and sum(e) <= order
        out[slot] = tuple(e[:dim])
    assert all(v is not None for v in out) and len(set(out)) == no
    return out


def _fact(e):
    out = 1.0
    for p in e:
        for t in range(2, p + 1):
            out *= t
    return out


def _poly_local(dxk, coef, exps):
    """sum_s coef[s] * prod_d dxk[..., d]^e_d / e_d!   (coef[s] IS the derivative value of slot s at the origin)"""
    f = torch.zeros(dxk.shape[:-1], dtype=torch.float64, device=dxk.device)
    for c, e in zip(coef, exps):
        term = torch.full_like(f, float(c) / _fact(e))
        for d, p in enumerate(e):
            if p:
                term = term * dxk[..., d] ** p
        f = f + term
    return f


def _report(got, exact_row, dim, order, label, factor=1.0):
    """got (n, no) on the device vs one row of exact derivative values; DOF-scaled, by derivative order"""
    exact = torch.as_tensor(exact_row, dtype=torch.float64, device=got.device)
    scale = exact.abs().clamp_min(1e-300)
    e = ((got - exact).abs() / scale)
    d = wl.dof_orders(dim, order)[: got.shape[1]]
    ok, lines = True, []
    for dd in range(order + 1):
        cols = torch.as_tensor(np.nonzero(d == dd)[0], device=got.device)
        ee = e[:, cols].flatten()
        if ee.numel() == 0:
            continue
        ee = ee[torch.isfinite(ee)]
        k99 = max(1, int(0.99 * ee.numel()))
        p99 = float(torch.kthvalue(ee, k99).values) if ee.numel() < 2 ** 24 else float(torch.quantile(ee[:: ee.numel() // 2 ** 23 + 1], 0.99))
        mx = float(ee.max())
        good = p99 <= factor * REPRO_P99 and mx <= factor * REPRO_MAX
        ok = ok and good
        lines.append(f"{label} d{dd}: p99 {p99:.2e} (tol {factor * REPRO_P99:.0e})  max {mx:.2e} (tol {factor * REPRO_MAX:.0e})"
                     f"  {'ok' if good else 'FAIL'}")
    print("\n".join(lines))
    assert ok, "\n".join(lines)


def _meta(n, k, order, knowns, wm):
    return (np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, knowns, np.int64), np.full(n, wm, np.int32))


@pytest.fixture(scope="module")
def cloud2d():
    """the headline cloud: 1M points 2D, k = 30 hoods from the device-side search (checked against cKDTree below)"""
    n, k = 1_000_000, 30
    x = wl.cloud(n, 2)
    x_d = torch.from_numpy(x).cuda()
    hoods_d = wlsqm.knn_hoods(x_d, k)
    return n, k, x, x_d, hoods_d


def test_full_size_hoods_bit_exact_vs_ckdtree(cloud2d):
    from scipy.spatial import cKDTree
    n, k, x, x_d, hoods_d = cloud2d
    sample = np.arange(0, n, 499)
    ref = cKDTree(x).query(x[sample], k + 1, workers=-1)[1][:, 1:]
    got = hoods_d[torch.from_numpy(sample).cuda()].cpu().numpy()
    assert np.array_equal(got, ref)


def test_cfg2_full_size_properties(cloud2d):
    n, k, x, x_d, hoods_d = cloud2d
    dim, order, no = 2, 4, 15
    exps = _slot_exps(dim, order)
    nk, od, kn, wm = _meta(n, k, order, 0, wlsqm.WEIGHT_UNIFORM)
    xk_d = wlsqm.gather(x_d, hoods_d)                           # (n, k, 2)
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=wlsqm.ALGO_BASIC, do_sens=False)
    s.prepare(x_d, xk_d)

    # ---- polynomial reproduction at every one of the 1M points: the data of case i is q(xk - xi) with one quartic q
    rng = np.random.default_rng(7)
    h = wl.H0
    coef = [rng.uniform(0.5, 1.5) * rng.choice([-1, 1]) / h ** sum(e) for e in exps]     # every term O(1) at |dx| ~ h
    fk_d = _poly_local(xk_d - x_d[:, None, :], coef, exps)
    fi_d = torch.zeros((n, no), dtype=torch.float64, device="cuda")
    assert s.solve(fk_d, fi_d) == 0
    _report(fi_d, coef, dim, order, "cfg2 1M polynomial reproduction")

    # ---- linearity in the data
    f = torch.from_numpy(wl.field(x)).cuda()
    g1 = wlsqm.gather(f, hoods_d)
    g2 = wlsqm.gather(torch.from_numpy(wl.field_step(wl.field(x), 3)).cuda(), hoods_d)
    r1, r2, r12 = (torch.zeros_like(fi_d) for _ in range(3))
    s.solve(g1, r1)
    s.solve(g2, r2)
    s.solve(2.0 * g1 - 3.0 * g2, r12)
    lin = 2.0 * r1 - 3.0 * r2
    sc = lin.abs().amax(dim=0).clamp_min(1e-300)
    err = ((r12 - lin).abs() / sc)
    d = wl.dof_orders(dim, order)
    for dd in range(order + 1):
        ee = err[:, torch.as_tensor(np.nonzero(d == dd)[0], device="cuda")]
        assert float(ee.max()) <= 100 * P99_TOL[dd], (dd, float(ee.max()))

    # ---- shard invariance: cases [n/2, n) on their own == the same rows of the whole batch, bit for bit
    lo = n // 2
    s2 = wlsqm.ExpertSolver(dim, nk[lo:], od[lo:], kn[lo:], wm[lo:], algorithm=wlsqm.ALGO_BASIC, do_sens=False)
    s2.prepare(x_d[lo:], xk_d[lo:])
    rs = torch.zeros((n - lo, no), dtype=torch.float64, device="cuda")
    s2.solve(g1[lo:], rs)
    assert torch.equal(rs, r1[lo:])
    del s2

    # ---- gather-on-device path == pre-gathered path, bit for bit
    s3 = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=wlsqm.ALGO_BASIC, do_sens=False)
    s3.prepare_hoods(x_d, hoods_d)
    rh = torch.zeros_like(fi_d)
    s3.solve_hoods(f, rh)
    assert torch.equal(rh, r1)
    del s3

    # ---- the oracle on every 997th case (1004 cases), noise-floor criterion
    idx = np.arange(0, n, 997)
    it = torch.from_numpy(idx).cuda()
    xk_s, fk_s, xi_s = xk_d[it].cpu().numpy(), g1[it].cpu().numpy(), x[idx]
    m = _meta(len(idx), k, order, 0, wlsqm.WEIGHT_UNIFORM)
    fi0 = np.zeros((len(idx), no))
    ref, _, _, _ = parity.oracle_solve(dim, *m, xi_s, xk_s, fk_s, fi0)
    a, b = parity.permuted_self_noise(dim, *m, xi_s, xk_s, fk_s, fi0)
    print(parity.check_against_floor(r1[it].cpu().numpy(), ref, b + (ref - a), dim, order, "cfg2 1M subsample vs oracle"))


def test_cfg5_full_size_interpolate(cloud2d):
    n, k, x, x_d, hoods_d = cloud2d
    dim, order, no = 2, 4, 15
    nk, od, kn, wm = _meta(n, k, order, 0, wlsqm.WEIGHT_UNIFORM)
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    s.prepare_hoods(x_d, hoods_d)
    f = torch.from_numpy(wl.field(x)).cuda()
    fi_d = torch.zeros((n, no), dtype=torch.float64, device="cuda")
    s.solve_hoods(f, fi_d)
    s.tree = object()       # I is given explicitly below
    nq = 16 * n
    I = torch.arange(n, device="cuda", dtype=torch.int64).repeat_interleave(16)
    g = torch.Generator(device="cuda").manual_seed(5)
    xq = x_d[I] + 0.3 * wl.H0 * (2 * torch.rand((nq, 2), dtype=torch.float64, device="cuda", generator=g) - 1)
    allv, _ = s.interpolate(xq, diff="all", I=I)
    assert tuple(allv.shape) == (nq, no)
    # the all-slots pass (in-place Taylor shift) and the reference-style per-slot calls (nested Horner form) are two
    # evaluation orders of the same polynomial: equal to a few ulp of the largest term
    for dslot in range(no):
        one, _ = s.interpolate(xq, diff=dslot, I=I)
        err = float((one - allv[:, dslot]).abs().max() / allv[:, dslot].abs().max())
        assert err <= 1e-13, (dslot, err)
    # evaluation at the model's own origin returns the stored coefficient exactly (every monomial but 1 vanishes)
    I0 = torch.arange(n, device="cuda", dtype=torch.int64)
    at0, _ = s.interpolate(x_d, diff="all", I=I0)
    assert torch.equal(at0, fi_d)
    # the model reproduces the data it interpolates: value at a query = Taylor sum (spot-check 4096 queries on the host)
    pick = torch.randint(0, nq, (4096,), device="cuda", generator=g)
    dx = (xq[pick] - x_d[I[pick]]).cpu().numpy()
    fi_h = fi_d[I[pick]].cpu().numpy()
    exps = _slot_exps(dim, order)
    val = sum(fi_h[:, sl] * dx[:, 0] ** e[0] * dx[:, 1] ** e[1] / _fact(e) for sl, e in enumerate(exps))
    got = allv[pick, 0].cpu().numpy()
    assert np.allclose(got, val, rtol=1e-12, atol=1e-14 * np.abs(val).max())


@pytest.mark.parametrize("dim,k,known_bdry", [(2, 24, "Y"), (1, 8, "X")])
def test_cfg4_full_size_mixed_knowns(dim, k, known_bdry):
    """2M points, order 3, WEIGHT_UNIFORM: F known in the interior, a first derivative known on every 1000th case"""
    n, order = 2_000_000, 3
    no = wlsqm.number_of_dofs(dim, order)
    exps = _slot_exps(dim, order)
    g = torch.Generator(device="cuda").manual_seed(11)
    h = wl.H0
    xi = h * n ** (1.0 / dim) * torch.rand((n, dim), dtype=torch.float64, device="cuda", generator=g)
    xk = xi[:, None, :] + 1.5 * h * (2 * torch.rand((n, k, dim), dtype=torch.float64, device="cuda", generator=g) - 1)
    bF = getattr(wlsqm, "b%d_F" % dim)
    bB = getattr(wlsqm, "b%d_%s" % (dim, known_bdry))
    iB = getattr(wlsqm, "i%d_%s" % (dim, known_bdry))
    nk, od, kn, wm = _meta(n, k, order, bF, wlsqm.WEIGHT_UNIFORM)
    kn[::1000] = bB
    rng = np.random.default_rng(3)
    coef = [rng.uniform(0.5, 1.5) * rng.choice([-1, 1]) / h ** sum(e) for e in exps]
    fk = _poly_local(xk - xi[:, None, :], coef, exps)
    fi = torch.full((n, no), 123.0, dtype=torch.float64, device="cuda")      # sentinel in every unknown slot
    fi[:, 0] = coef[0]
    fi[::1000, 0] = 123.0
    fi[::1000, iB] = coef[iB]
    fi_in = fi.clone()
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    if dim == 1:
        s.prepare(xi[:, 0].contiguous(), xk[:, :, 0].contiguous())
    else:
        s.prepare(xi, xk)
    s.solve(fk, fi)
    # knowns untouched, bit for bit, at every point
    interior = torch.ones(n, dtype=torch.bool, device="cuda")
    interior[::1000] = False
    assert torch.equal(fi[interior, 0], fi_in[interior, 0])
    assert torch.equal(fi[~interior, iB], fi_in[~interior, iB])
    _report(fi, coef, dim, order, f"cfg4 {dim}D 2M mixed knowns", factor=10.0 if dim == 1 else 1.0)


def test_cfg3_shard_size_iterative_sens():
    """config 3 at the size of one rank's shard on 8 GPUs (500k of 4M points): 3D order 4, k = 60, F known,
    ALGO_ITERATIVE max_iter = 3, do_sens"""
    n, dim, order, k = 500_000, 3, 4, 60
    no = 35
    exps = _slot_exps(dim, order)
    g = torch.Generator(device="cuda").manual_seed(13)
    h = wl.H0
    xi = h * n ** (1.0 / dim) * torch.rand((n, dim), dtype=torch.float64, device="cuda", generator=g)
    xk = xi[:, None, :] + 1.5 * h * (2 * torch.rand((n, k, dim), dtype=torch.float64, device="cuda", generator=g) - 1)
    nk, od, kn, wm = _meta(n, k, order, wlsqm.b3_F, wlsqm.WEIGHT_CENTER)
    rng = np.random.default_rng(5)
    coef = [rng.uniform(0.5, 1.5) * rng.choice([-1, 1]) / h ** sum(e) for e in exps]
    fk = _poly_local(xk - xi[:, None, :], coef, exps)
    fi = torch.zeros((n, no), dtype=torch.float64, device="cuda")
    fi[:, 0] = coef[0]
    sens = torch.zeros((n, k, no), dtype=torch.float64, device="cuda")
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=wlsqm.ALGO_ITERATIVE, do_sens=True, max_iter=3)
    s.prepare(xi, xk)
    iters = s.solve(fk, fi, sens)
    assert 1 <= iters <= 3
    assert torch.equal(fi[:, 0], torch.full((n,), coef[0], dtype=torch.float64, device="cuda"))
    _report(fi, coef, dim, order, "cfg3 500k polynomial reproduction (iterative)")
    # sens: NaN exactly in the known slot; sens^T (fk - F) reproduces the unknown DOFs (the known F enters through
    # the elimination, impl.pyx:792-818: with the data shifted by F the known value is 0 and contributes nothing)
    nanpat = torch.isnan(sens)
    assert bool(nanpat[:, :, 0].all()) and not bool(nanpat[:, :, 1:].any())
    lin = torch.einsum("ikj,ik->ij", sens[:, :, 1:], fk - coef[0])
    exact = torch.as_tensor(coef[1:], dtype=torch.float64, device="cuda")
    e = ((lin - exact).abs() / exact.abs())
    d = wl.dof_orders(dim, order)[1:]
    for dd in range(1, order + 1):
        ee = e[:, torch.as_tensor(np.nonzero(d == dd)[0], device="cuda")]
        assert float(ee.max()) <= REPRO_MAX, (dd, float(ee.max()))


# ---------------------------------------------------------------------------------------------------------------
# configs 3 and 4 at their FULL sizes on real kNN neighbourhoods, with an oracle subsample
# ---------------------------------------------------------------------------------------------------------------
def _subsample_vs_oracle(dim, idx, meta_all, x_d, hoods_d, f_d, fi_in_d, fi_out_d, algorithm, max_iter, label,
                         sens_d=None, iters=None):
    """the oracle on the cases `idx` (host), against the rows of the full-size GPU result: noise floor of
    tests/parity.py per order group; returns the report"""
    it = torch.from_numpy(idx).cuda()
    hd = hoods_d[it].long()
    x2 = x_d if x_d.dim() == 2 else x_d[:, None]
    xk_s = x2[hd].cpu().numpy()
    if dim == 1:
        xk_s = np.ascontiguousarray(xk_s[:, :, 0])
    fk_s = f_d[hd].cpu().numpy()
    xi_s = x_d[it].cpu().numpy()
    nk, od, kn, wm = (a[idx] for a in meta_all)
    fi0 = fi_in_d[it].cpu().numpy()
    ref, sens_o, _, so = parity.oracle_solve(dim, nk, od, kn, wm, xi_s, xk_s, fk_s, fi0, algorithm, sens_d is not None, max_iter)
    a, b = parity.hetero_self_noise(dim, nk, od, kn, wm, xi_s, xk_s, fk_s, fi0, algorithm, max_iter)
    got = fi_out_d[it].cpu().numpy()
    order = int(od[0])
    rep = []
    for knv in np.unique(kn):          # one group per knowns pattern (interior / boundary)
        m = kn == knv
        rep.append(parity.check_against_floor(got[m], ref[m], (b + (ref - a))[m], dim, order,
                                              "%s knowns=%d (%d cases)" % (label, int(knv), int(m.sum()))))
    if sens_d is not None:
        parity.check_sens(sens_d[it].cpu().numpy(), sens_o, label)
    if iters is not None:
        # (per-case counts may differ by a round where the bit-exact `norm == prev_norm` exit fires on one side only)
        assert iters[idx].max() == so.iters.max() and iters[idx].min() >= 0
        rep.append("%s: per-case refinement iterations gpu == oracle for %d of %d cases" % (label, int((iters[idx] == so.iters).sum()), len(idx)))
    return "\n".join(rep)


@pytest.mark.parametrize("dim,k", [(2, 24), (1, 8)])
def test_cfg4_full_size_knn_hoods_boundary_construction(dim, k):
    """BASELINE.json configs[3] as SURVEY.md 8d builds it: 2M points, order 3, WEIGHT_UNIFORM; interior points know F;
    boundary points (2D: the 4 sqrt(n) points nearest the box edges; 1D: the two end points and every 1000th point) know
    the derivative dF/dy (dF/dx in 1D) from the analytic field instead, and take their neighbours from the interior
    points only (the Neumann-boundary stencil of wlsqm/fitter/defs.pyx:199-207).  Neighbourhoods are real k-nearest-
    neighbour lists (device-side search, spot-checked against cKDTree)."""
    from scipy.spatial import cKDTree
    n, order = 2_000_000, 3
    no = wlsqm.number_of_dofs(dim, order)
    p = parity.cfg4_problem(dim, n, k)
    x, x2, bnd, interior, boundary, hoods_d, iB = (p[key] for key in ("x", "x2", "bnd", "interior", "boundary", "hoods_d", "iB"))
    nk, od, kn, wm = p["meta"]
    x_d, f_d, fi_in_d = p["x_d"], p["f_d"], p["fi_in_d"]
    # hood indexing: bit exact against cKDTree on a sample of both groups
    tree = cKDTree(x2[interior])
    si = interior[:: max(1, len(interior) // 500)]
    ref_i = interior[tree.query(x2[si], k + 1)[1][:, 1:]]
    assert np.array_equal(hoods_d[torch.from_numpy(si).cuda()].cpu().numpy(), ref_i)
    sb = boundary[:: max(1, len(boundary) // 300)]
    ref_b = interior[tree.query(x2[sb], k)[1]]
    assert np.array_equal(hoods_d[torch.from_numpy(sb).cuda()].cpu().numpy(), ref_b)
    fi_d = fi_in_d.clone()
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    s.prepare_hoods(x_d, hoods_d)
    assert s.solve_hoods(f_d, fi_d) == 0
    # knowns untouched bit for bit at every point; every unknown slot overwritten
    bd = torch.from_numpy(bnd).cuda()
    assert torch.equal(fi_d[~bd, 0], fi_in_d[~bd, 0]) and torch.equal(fi_d[bd, iB], fi_in_d[bd, iB])
    unk = torch.ones((n, no), dtype=torch.bool, device="cuda")
    unk[~bd, 0] = False
    unk[bd, iB] = False
    assert not bool((fi_d[unk] == 123.0).any())
    # sanity: the fitted value at the boundary points against the analytic field (one-sided order-3 stencils of radius
    # ~5 h0: truncation error ~ (pi r)^4 / 24)
    assert float((fi_d[bd, 0] - f_d[bd]).abs().max()) < 1e-3
    # the oracle on ~1000 interior and ~500 boundary cases, noise floor per group
    idx = np.sort(np.concatenate([interior[:: len(interior) // 1000], boundary[:: max(1, len(boundary) // 500)]]))
    print(_subsample_vs_oracle(dim, idx, (nk, od, kn, wm), x_d, hoods_d, f_d, fi_in_d, fi_d, 1, 0,
                               "cfg4 %dD 2M kNN hoods" % dim))


def test_cfg3_full_size_4M_knn_hoods_iterative_sens():
    """BASELINE.json configs[2] at its FULL size on one GPU: 4M points 3D, order 4, k = 60 nearest neighbours, F known,
    ALGO_ITERATIVE max_iter = 3, do_sens (about 150 GB of the 180 GB: operators 66 GB + sens 67 GB + geometry).
    Checked: knowns bit for bit, NaN pattern of sens at every point, sens^T-identity on a sample, shard invariance
    (one rank's 500k-point range of the 8-GPU split solved alone == the same rows), and the oracle on ~1000 cases
    (noise floor, sens, per-case iteration counts).  Skipped when the device has less than 160 GB free."""
    free, total = torch.cuda.mem_get_info()
    torch.cuda.empty_cache()
    wlsqm.pool_trim()
    free, total = torch.cuda.mem_get_info()
    if free < 160e9:
        pytest.skip("needs 160 GB of free device memory, have %.0f GB" % (free / 1e9))
    n, dim, order, k, no = 4_000_000, 3, 4, 60, 35
    x = wl.cloud(n, dim)
    x_d = torch.from_numpy(x).cuda()
    hoods_d = wlsqm.knn_hoods(x_d, k)
    from scipy.spatial import cKDTree
    sample = np.arange(0, n, 7993)
    ref_h = cKDTree(x).query(x[sample], k + 1, workers=-1)[1][:, 1:]
    assert np.array_equal(hoods_d[torch.from_numpy(sample).cuda()].cpu().numpy(), ref_h)
    nk, od, kn, wm = _meta(n, k, order, wlsqm.b3_F, wlsqm.WEIGHT_CENTER)
    f = wl.field(x)
    f_d = torch.from_numpy(f).cuda()
    fi_in_d = torch.zeros((n, no), dtype=torch.float64, device="cuda")
    fi_in_d[:, 0] = f_d
    fi_d = fi_in_d.clone()
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=wlsqm.ALGO_ITERATIVE, do_sens=True, max_iter=3)
    s.prepare_hoods(x_d, hoods_d)
    sens = torch.empty((n, k, no), dtype=torch.float64, device="cuda")
    iters = s.solve_hoods(f_d, fi_d, sens)
    assert 1 <= iters <= 3
    assert torch.equal(fi_d[:, 0], f_d)
    # NaN exactly in the known slot, at every point (chunked: the mask of the whole array would be 8 GB)
    for c0 in range(0, n, 500_000):
        blk = torch.isnan(sens[c0:c0 + 500_000])
        assert bool(blk[:, :, 0].all()) and not bool(blk[:, :, 1:].any())
        del blk
    its = s.iterations()
    # the oracle on every 3989th case
    idx = np.arange(0, n, 3989)
    print(_subsample_vs_oracle(dim, idx, (nk, od, kn, wm), x_d, hoods_d, f_d, fi_in_d, fi_d, 2, 3, "cfg3 4M kNN hoods",
                               sens_d=sens, iters=its))
    # shard invariance: rank 5 of 8 owns [2.5M, 3M); solved alone it gives the same rows, bit for bit
    keep_rows = fi_d[2_500_000:3_000_000].clone()
    keep_sens = sens[2_500_000:2_500_100].clone()
    del sens, s
    torch.cuda.empty_cache()
    wlsqm.pool_trim()
    lo, hi = 2_500_000, 3_000_000
    s2 = wlsqm.ExpertSolver(dim, nk[lo:hi], od[lo:hi], kn[lo:hi], wm[lo:hi], algorithm=wlsqm.ALGO_ITERATIVE, do_sens=True,
                            max_iter=3)
    s2.prepare_hoods(x_d, hoods_d[lo:hi], xi=x_d[lo:hi])
    fi2 = fi_in_d[lo:hi].clone()
    sens2 = torch.empty((hi - lo, k, no), dtype=torch.float64, device="cuda")
    s2.solve_hoods(f_d, fi2, sens2)
    assert torch.equal(fi2, keep_rows)
    assert torch.equal(torch.nan_to_num(sens2[:100]), torch.nan_to_num(keep_sens))
    del s2, sens2
    torch.cuda.empty_cache()
    wlsqm.pool_trim()
