"""GPU parity tests: the CUDA path, through the C ABI (wlsqm_b200 -> libwlsqm_b200.so), vs the oracle.

Tolerances: known-answer tests use the reference's own test tolerances (tests/test_simple.py:24
ATOL_EXACT = 1e-10, tests/test_expert.py 1e-12/1e-14 equivalences are replaced by oracle comparisons);
differential tests use the per-derivative-order noise-floor criterion of tests/parity.py
(SURVEY.md 8c): <= 4x the oracle's own neighbour-permutation self-difference, p99 <= 1e-10 for d <= 1.
"""
import numpy as np
import pytest

import parity
import oracle as orc

pytestmark = pytest.mark.gpu

wlsqm = pytest.importorskip("wlsqm_b200")


def _uniform_meta(n, k, order, knowns, wm):
    return (np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, knowns, np.int64),
            np.full(n, wm, np.int32))


def _run_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0, algorithm=1, do_sens=False, max_iter=10, debug=False):
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=algorithm, do_sens=do_sens, max_iter=max_iter,
                           ntasks=1, debug=debug)
    s.prepare(x, xk)
    fi = fi0.copy()
    sens = np.zeros((len(nk), xk.shape[1], fi.shape[1])) if do_sens else None
    it = s.solve(fk, fi, sens)
    return fi, sens, it, s


CONFIGS = [
    # dim, order, k, knowns, wm, algorithm, n, do_sens
    pytest.param(2, 2, 12, 1, 2, 1, 4000, False, id="cfg1-2D-o2-k12-bF-center"),
    pytest.param(2, 4, 30, 0, 1, 1, 4000, False, id="cfg2-2D-o4-k30-uniform"),
    pytest.param(2, 4, 30, 0, 2, 1, 4000, False, id="cfg2-2D-o4-k30-center"),
    pytest.param(2, 4, 30, 1, 1, 1, 4000, True, id="cfg2v-2D-o4-k30-bF-sens"),
    pytest.param(2, 4, 30, 1, 1, 2, 3000, False, id="cfg2-2D-o4-k30-bF-iterative"),
    pytest.param(3, 4, 60, 1, 2, 2, 800, True, id="cfg3-3D-o4-k60-bF-iter-sens"),
    pytest.param(3, 4, 60, 0, 1, 1, 600, False, id="3D-o4-k60-nr35"),
    pytest.param(3, 2, 20, 1, 2, 1, 2000, False, id="3D-o2-k20"),
    pytest.param(3, 3, 40, 0, 1, 1, 1000, False, id="3D-o3-k40"),
    pytest.param(1, 3, 8, 1, 1, 1, 4000, False, id="cfg4-1D-o3-k8"),
    pytest.param(2, 3, 24, 1, 1, 1, 4000, False, id="cfg4-2D-o3-k24"),
    pytest.param(1, 4, 9, 0, 2, 2, 2000, True, id="1D-o4-k9-iter-sens"),
    pytest.param(2, 1, 6, 0, 2, 1, 2000, False, id="2D-o1-k6"),
    pytest.param(2, 0, 5, 0, 1, 1, 1000, False, id="2D-o0-k5"),
]


@pytest.mark.parametrize("dim,order,k,knowns,wm,algo,n,do_sens", CONFIGS)
def test_differential_vs_oracle(dim, order, k, knowns, wm, algo, n, do_sens):
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    no = wlsqm.number_of_dofs(dim, order)
    nk, od, kn, w = _uniform_meta(n, k, order, knowns, wm)
    fi0 = np.zeros((n, no))
    fi0[:, 0] = f
    max_iter = 3
    fi_g, sens_g, it_g, s = _run_gpu(dim, nk, od, kn, w, x, xk, fk, fi0, algo, do_sens, max_iter)
    fi_o, sens_o, it_o, _ = parity.oracle_solve(dim, nk, od, kn, w, x, xk, fk, fi0, algo, do_sens, max_iter)
    a, b = parity.permuted_self_noise(dim, nk, od, kn, w, x, xk, fk, fi0, algo, max_iter)
    print(parity.check_against_floor(fi_g, fi_o, b + (fi_o - a), dim, order, "gpu-vs-oracle"))
    assert it_g == it_o
    # knowns are left untouched, bit for bit
    for o in range(no):
        if knowns >> o & 1:
            assert np.array_equal(fi_g[:, o], fi0[:, o])
    if do_sens:
        assert np.array_equal(np.isnan(sens_g), np.isnan(sens_o))
        sc = np.nanmax(np.abs(np.where(np.isnan(sens_o), 0.0, sens_o)), axis=(0, 1))
        sc[sc == 0] = 1.0
        err = np.abs(np.nan_to_num(sens_g) - np.nan_to_num(sens_o)) / sc
        # sens entries are solves of the same scaled LU: cond * eps relative to the largest entry of the column
        assert err.max() < 1e-8, err.max()
        assert np.median(err) < 1e-12


@pytest.mark.parametrize("dim,bucketed", [(2, False), (2, True), (3, True)])
def test_heterogeneous_batch(dim, bucketed, monkeypatch):
    """per-case nk / order / knowns / weighting in one batch (expert.pyx:92-104); `bucketed`: prepare() runs one launch
    per order over that order's case list (the default from 8192 cases on) instead of the maximum order's kernel for all.
    Criterion: the noise floor of tests/parity.py per ORDER GROUP (the oracle against itself with every case's own
    neighbours permuted), against the oracle and against the unmodified reference's outputs."""
    monkeypatch.setenv("WLSQM_PREP_BUCKET_MIN", "1000" if bucketed else "1000000000")
    c = parity.hetero_case(dim)
    n, x, xk, fk, nk, od, kn, wm, fi0 = (c[k] for k in ("n", "x", "xk", "fk", "nk", "od", "kn", "wm", "fi0"))
    fi_g, sens_g, _, _ = _run_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0, 1, True)
    fi_o, sens_o, _, _ = parity.oracle_solve(dim, nk, od, kn, wm, x, xk, fk, fi0, 1, True)
    a, b = parity.hetero_self_noise(dim, nk, od, kn, wm, x, xk, fk, fi0, 1, seeds=(7, 8, 9))
    # untouched: columns >= no_j, and known slots
    for j in range(n):
        no = wlsqm.number_of_dofs(dim, int(od[j]))
        assert np.array_equal(fi_g[j, no:], fi0[j, no:])
        for o in range(no):
            if kn[j] >> o & 1:
                assert fi_g[j, o] == fi0[j, o]
    print(parity.check_hetero_against_floor(fi_g, fi_o, b + (fi_o - a), dim, od, "hetero gpu-vs-oracle"))
    assert np.array_equal(np.isnan(sens_g), np.isnan(sens_o))
    parity.check_sens(sens_g, sens_o, "hetero")
    # sens rows k >= nk_j and columns o >= no_j stay untouched (zeros here)
    for j in range(0, n, 97):
        assert (sens_g[j, nk[j]:, :] == 0).all()
    # the unmodified reference on the same inputs (tests/golden/make_golden_hetero.py)
    g = parity.golden_hetero(dim)
    assert g is not None, "tests/golden/golden_hetero.npz is missing"
    print(parity.check_hetero_against_floor(fi_g, g["fi_ref"], b + (g["fi_ref"] - a), dim, od, "hetero gpu-vs-reference"))
    assert np.array_equal(np.packbits(np.isnan(sens_g).ravel()), g["sens_nan_bits"])
    parity.check_sens(sens_g[::50], g["sens_ref_every50"], "hetero vs reference")
    if bucketed:
        # same arithmetic per fit in the per-order kernels as in the maximum order's: compare the two launch plans
        monkeypatch.setenv("WLSQM_PREP_BUCKETS", "0")
        fi_u, sens_u, _, _ = _run_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0, 1, True)
        print(parity.check_hetero_against_floor(fi_u, fi_o, b + (fi_o - a), dim, od, "hetero one-launch-vs-oracle"))
        sc = np.abs(fi_u).max(axis=0)
        sc[sc == 0] = 1
        d = np.abs(fi_g - fi_u) / sc
        print("per-order launches vs one launch: max scaled difference %.3g, identical: %s" % (d.max(), np.array_equal(fi_g, fi_u)))
        assert np.array_equal(np.isnan(sens_g), np.isnan(sens_u))


@pytest.mark.parametrize("dim,algo,seed", [(1, 1, 11), (1, 2, 12), (2, 1, 13), (2, 2, 14), (3, 1, 15), (3, 2, 16)])
def test_random_knowns_masks_orders_and_sizes(dim, algo, seed):
    """every case draws its own order, ARBITRARY knowns bitmask (any subset of its DOFs, including none and all but
    one -- the reference's remap(), infra.pyx:145-200, not only the b*_F / b*_XY patterns of its tests), weighting and
    neighbour count; both algorithms, all three dimensions.  Criterion: the per-order-group noise floor of
    tests/parity.py; untouched entries (known slots, columns past no_j) stay bit-for-bit."""
    n = 1500 if dim == 3 else 2500
    kmax = {1: 12, 2: 30, 3: 60}[dim]
    nomax = wlsqm.number_of_dofs(dim, 4)
    x, hoods, f = parity.make_case(n, dim, kmax, seed=seed)
    xk, fk = parity.gathered(x, f, hoods)
    rng = np.random.default_rng(seed)
    od = rng.integers(0, 5, n).astype(np.int32)
    no = np.array([wlsqm.number_of_dofs(dim, int(o)) for o in od])
    kn = np.zeros(n, np.int64)
    for j in range(n):
        bits = rng.random(no[j]) < rng.choice([0.0, 0.15, 0.5, 0.9])
        if bits.all():
            bits[rng.integers(no[j])] = False                     # (at least one unknown)
        kn[j] = int(sum(1 << o for o in range(no[j]) if bits[o]))
    nr = np.array([no[j] - bin(int(kn[j])).count("1") for j in range(n)])
    lo = np.minimum(kmax, np.maximum(nr + 2, (3 * nr) // 2 + 1))
    nk = rng.integers(lo, kmax + 1).astype(np.int32)
    wm = rng.integers(1, 3, n).astype(np.int32)
    # known slots hold the exact derivatives of the sampled field where the model has them (else a fit with made-up
    # "known" derivatives is far from the data and only tests conditioning), other slots random
    fi0 = rng.standard_normal((n, nomax))
    exact, _, _, _ = parity.oracle_solve(dim, np.full(n, kmax, np.int32), np.full(n, 4, np.int32), np.zeros(n, np.int64),
                                         np.ones(n, np.int32), x, xk, fk, np.zeros((n, nomax)), 1)
    for j in range(n):
        for o in range(no[j]):
            if kn[j] >> o & 1:
                fi0[j, o] = exact[j, o]
    fi_g, sens_g, it_g, _ = _run_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0, algo, True, max_iter=6)
    fi_o, sens_o, it_o, _ = parity.oracle_solve(dim, nk, od, kn, wm, x, xk, fk, fi0, algo, True, max_iter=6)
    # (the maximum over a few hundred cases is a heavy-tailed statistic -- one case leaving its refinement loop a round
    # earlier under another summation order moves it by 5x: eight permutations for the floor)
    a, b = parity.hetero_self_noise(dim, nk, od, kn, wm, x, xk, fk, fi0, algo, max_iter=6, seeds=tuple(range(7, 15)),
                                    iter_rounds=True)

    def untouched(fi):
        for j in range(n):
            assert np.array_equal(fi[j, no[j]:], fi0[j, no[j]:])
            for o in range(no[j]):
                if kn[j] >> o & 1:
                    assert fi[j, o] == fi0[j, o]
    untouched(fi_g)
    print(parity.check_hetero_against_floor(fi_g, fi_o, b + (fi_o - a), dim, od, "random masks %dD algo %d" % (dim, algo)))
    parity.check_sens(sens_g, sens_o, "random masks %dD algo %d" % (dim, algo))
    if algo == 2:
        assert it_g == it_o, (it_g, it_o)
    # the one-shot entry points on the same batch (simple.pyx fit_*D_many / fit_*D_iterative_many: the fused kernel for
    # ALGO_BASIC without sens, which eliminates the knowns in its own way)
    fi_1 = fi0.copy()
    if algo == 1:
        getattr(wlsqm, "fit_%dD_many_parallel" % dim)(xk, fk, nk, x, fi_1, None, 0, od, kn, wm)
    else:
        getattr(wlsqm, "fit_%dD_iterative_many" % dim)(xk, fk, nk, x, fi_1, None, 0, od, kn, wm, max_iter=6)
    untouched(fi_1)
    print(parity.check_hetero_against_floor(fi_1, fi_o, b + (fi_o - a), dim, od, "random masks, one-shot %dD algo %d" % (dim, algo)))
    # the fitted models evaluated at random points: every slot in one pass and a few single slots (expert.pyx:687-781;
    # per-model orders), same coefficients on both sides
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=algo, max_iter=6, ntasks=1)
    s.prepare(x, xk)
    fi_s = fi0.copy()
    s.solve(fk, fi_s)
    s.prep_interpolate()
    nq = 1501
    I = rng.integers(0, n, nq).astype(np.int64)
    xq = x[I] + 2e-3 * rng.uniform(-1, 1, x[I].shape)
    so = orc.OracleSolver(dim, nk, od, kn, wm)
    so.xi = x
    so.fi = fi_s.copy()
    out_all, _ = s.interpolate(xq, diff='all', I=I)
    for d in sorted(set([0, nomax - 1] + rng.integers(0, nomax, 3).tolist())):
        og, _ = s.interpolate(xq, diff=d, I=I)
        oo = so.interpolate(xq, I, d)
        assert np.array_equal(np.isnan(og), np.isnan(oo)), d
        assert np.nanmax(np.abs(og - oo), initial=0.0) <= 1e-12 * max(np.nanmax(np.abs(oo), initial=0.0), 1e-300), d
        assert np.nanmax(np.abs(out_all[:, d] - oo), initial=0.0) <= 1e-12 * max(np.nanmax(np.abs(oo), initial=0.0), 1e-300), d


@pytest.mark.parametrize("dim,order,k,nkn,algo,seed", [
    (1, 4, 9, 1, 1, 21), (1, 3, 8, 2, 1, 22), (2, 3, 24, 1, 1, 23), (2, 4, 30, 3, 1, 24), (2, 2, 12, 2, 1, 25),
    (3, 2, 20, 2, 1, 26), (3, 4, 60, 33, 1, 27), (3, 3, 40, 4, 2, 28), (2, 4, 30, 2, 2, 29), (3, 4, 60, 2, 2, 30),
    (3, 4, 45, 0, 2, 31), (3, 4, 61, 1, 2, 32)])
def test_random_knowns_patterns_of_equal_count(dim, order, k, nkn, algo, seed):
    """same sizes everywhere, but every case knows a DIFFERENT set of nkn DOFs: the geometry-uniform kernels (one size
    record by value, one 8-byte mask per case; the packed solve for the small models) against the oracle"""
    n = 1200 if dim == 3 else 3000
    no = wlsqm.number_of_dofs(dim, order)
    x, hoods, f = parity.make_case(n, dim, k, seed=seed)
    xk, fk = parity.gathered(x, f, hoods)
    rng = np.random.default_rng(seed)
    kn = np.array([int(sum(1 << int(o) for o in rng.choice(no, nkn, replace=False))) for _ in range(n)], np.int64)
    nk, od, _, wm = _uniform_meta(n, k, order, 0, 2)
    exact, _, _, _ = parity.oracle_solve(dim, nk, od, np.zeros(n, np.int64), wm, x, xk, fk, np.zeros((n, no)), 1)
    fi0 = np.where((kn[:, None] >> np.arange(no)[None, :]) & 1, exact, rng.standard_normal((n, no)))
    fi_g, sens_g, it_g, _ = _run_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0, algo, True, max_iter=5)
    fi_o, sens_o, it_o, _ = parity.oracle_solve(dim, nk, od, kn, wm, x, xk, fk, fi0, algo, True, max_iter=5)
    a, b = parity.hetero_self_noise(dim, nk, od, kn, wm, x, xk, fk, fi0, algo, max_iter=5, seeds=tuple(range(7, 13)),
                                    iter_rounds=True)
    known = ((kn[:, None] >> np.arange(no)[None, :]) & 1).astype(bool)
    assert np.array_equal(fi_g[known], fi0[known])
    print(parity.check_against_floor(fi_g, fi_o, b + (fi_o - a), dim, order, "equal-count masks %dD o%d" % (dim, order)))
    parity.check_sens(sens_g, sens_o, "equal-count masks")
    if algo == 2:
        assert it_g == it_o
    # CUDA tensors (no staging; strided fk view)
    torch = pytest.importorskip("torch")
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=algo, max_iter=5)
    s.prepare(torch.from_numpy(x).cuda(), torch.from_numpy(xk).cuda())
    fi_d = torch.from_numpy(fi0).cuda()
    s.solve(torch.from_numpy(fk).cuda(), fi_d)
    assert np.array_equal(fi_d.cpu().numpy(), fi_g)


@pytest.mark.parametrize("dim,do_sens", [(2, True), (2, False), (3, True), (3, False)])
def test_heterogeneous_batch_iterative(dim, do_sens):
    """per-case records AND ALGO_ITERATIVE (solve_kernel<DIM, ITER = true, SENS, UNI = false>): the in-kernel refinement
    with the bit-pattern-dependent `norm == prev_norm` exit (impl.pyx:1057-1060) on cases of mixed sizes.  Against the
    oracle (noise floor per order group, per-case iteration counts equal) and against the unmodified reference's
    outputs for the same inputs (tests/golden/golden_hetero_iter.npz: fi, returned iteration count, sens)."""
    c = parity.hetero_case(dim)
    n, x, xk, fk, nk, od, kn, wm, fi0 = (c[k] for k in ("n", "x", "xk", "fk", "nk", "od", "kn", "wm", "fi0"))
    fi_g, sens_g, it_g, s = _run_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0, 2, do_sens, 3)
    fi_o, sens_o, it_o, so = parity.oracle_solve(dim, nk, od, kn, wm, x, xk, fk, fi0, 2, do_sens, 3)
    a, b = parity.hetero_self_noise(dim, nk, od, kn, wm, x, xk, fk, fi0, 2, 3, seeds=(7, 8, 9))
    assert it_g == it_o
    # per case, the exit `norm == prev_norm` (impl.pyx:1057-1060) compares bit patterns: where two successive residual
    # norms agree to the last bit on one side only, the counts differ by a round ("document any mismatch");
    # what the reference returns -- the maximum -- is equal (asserted above), every count is a legal one
    its = s.iterations()
    assert its.min() >= 0 and its.max() <= 3 and its.max() == it_g
    print("per-case refinement iterations: gpu == oracle for %d of %d cases" % (int((its == so.iters).sum()), n))
    for j in range(0, n, 7):
        no = wlsqm.number_of_dofs(dim, int(od[j]))
        assert np.array_equal(fi_g[j, no:], fi0[j, no:])
        for o in range(no):
            if kn[j] >> o & 1:
                assert fi_g[j, o] == fi0[j, o]
    print(parity.check_hetero_against_floor(fi_g, fi_o, b + (fi_o - a), dim, od, "hetero-iterative gpu-vs-oracle"))
    g = parity.golden_hetero_iter(dim)
    assert g is not None, "tests/golden/golden_hetero_iter.npz is missing"
    assert it_g == int(g["iters_max"][0])
    print(parity.check_hetero_against_floor(fi_g, g["fi_ref"], b + (g["fi_ref"] - a), dim, od, "hetero-iterative gpu-vs-reference"))
    if do_sens:
        assert np.array_equal(np.isnan(sens_g), np.isnan(sens_o))
        parity.check_sens(sens_g, sens_o, "hetero-iterative")
        assert np.array_equal(np.packbits(np.isnan(sens_g).ravel()), g["sens_nan_bits"])
        parity.check_sens(sens_g[::50], g["sens_ref_every50"], "hetero-iterative vs reference")


def test_fk_alias_of_fi_on_device():
    """fk given as a view into fi (expert.pyx:548-557): every case must see the old data"""
    torch = pytest.importorskip("torch")
    n, k = 2000, 12
    x, hoods, f = parity.make_case(n, 2, k)
    nk, od, kn, w = _uniform_meta(n, k, 2, 0, 2)
    xk, fk = parity.gathered(x, f, hoods)
    fi_o = np.zeros((n, 6))
    parity_fi, _, _, _ = parity.oracle_solve(2, nk, od, kn, w, x, xk, fk, fi_o)
    # device buffer holding [fk | fi] side by side so that fk is a strided view of the same allocation
    buf = torch.zeros((n, k + 6), dtype=torch.float64, device="cuda")
    buf[:, :k] = torch.from_numpy(fk).cuda()
    s = wlsqm.ExpertSolver(2, nk, od, kn, w)
    s.prepare(torch.from_numpy(x).cuda(), torch.from_numpy(xk).cuda())
    s.solve(buf[:, :k], buf[:, k:])
    torch.cuda.synchronize()
    got = buf[:, k:].cpu().numpy()
    sc = np.abs(parity_fi).max(axis=0)
    assert (np.abs(got - parity_fi) / sc).max() < 1e-9


def test_torch_tensors_zero_copy_matches_numpy_path():
    torch = pytest.importorskip("torch")
    n, k = 3000, 30
    x, hoods, f = parity.make_case(n, 2, k)
    nk, od, kn, w = _uniform_meta(n, k, 4, 1, 1)
    xk, fk = parity.gathered(x, f, hoods)
    fi0 = np.zeros((n, 15))
    fi0[:, 0] = f
    fi_np, _, _, _ = _run_gpu(2, nk, od, kn, w, x, xk, fk, fi0)
    s = wlsqm.ExpertSolver(2, nk, od, kn, w)
    s.prepare(torch.from_numpy(x).cuda(), torch.from_numpy(xk).cuda())
    fi_t = torch.from_numpy(fi0).cuda()
    s.solve(torch.from_numpy(fk).cuda(), fi_t)
    torch.cuda.synchronize()
    assert np.array_equal(fi_t.cpu().numpy(), fi_np)


def test_conds_match_svd_of_oracle_scaled_matrix():
    n, k = 500, 30
    x, hoods, f = parity.make_case(n, 2, k)
    nk, od, kn, w = _uniform_meta(n, k, 4, 0, 1)
    xk, fk = parity.gathered(x, f, hoods)
    fi0 = np.zeros((n, 15))
    _, _, _, s = _run_gpu(2, nk, od, kn, w, x, xk, fk, fi0, debug=True)
    _, _, _, so = parity.oracle_solve(2, nk, od, kn, w, x, xk, fk, fi0)
    cg, co = s.conds(), so.conds()
    assert np.allclose(cg, co, rtol=1e-6), np.abs(cg / co - 1).max()


def test_interpolate_nearest_all_diffs_vs_oracle():
    n, k, dim, order = 3000, 30, 2, 4
    x, hoods, f = parity.make_case(n, dim, k)
    nk, od, kn, w = _uniform_meta(n, k, order, 0, 2)
    xk, fk = parity.gathered(x, f, hoods)
    fi0 = np.zeros((n, 15))
    fi_g, _, _, s = _run_gpu(dim, nk, od, kn, w, x, xk, fk, fi0)
    rng = np.random.default_rng(5)
    xq = np.repeat(x, 4, axis=0) + 0.3e-2 * rng.uniform(-1, 1, (4 * n, 2))
    s.prep_interpolate()
    out0, I = s.interpolate(xq, mode='nearest', diff=0)
    from scipy.spatial import cKDTree
    assert np.array_equal(I, cKDTree(x).query(xq)[1])
    assert I.dtype == np.int_
    so = orc.OracleSolver(dim, nk, od, kn, w)
    so.xi = x
    so.fi = fi_g.copy()      # same coefficients: isolates the evaluator
    allg, _ = s.interpolate(xq, diff='all', I=I)
    for d in range(15):
        og, I2 = s.interpolate(xq, mode='nearest', diff=d, I=I)
        oo = so.interpolate(xq, I, d)
        scale = max(np.abs(oo).max(), 1e-300)
        assert np.abs(og - oo).max() / scale < 1e-12, (d, np.abs(og - oo).max() / scale)
        # all-slots pass (in-place Taylor shift) vs per-slot call (nested Horner): two evaluation orders
        assert np.abs(allg[:, d] - og).max() <= 1e-13 * scale
    # a model index equal to ncases ("no neighbour found") poisons the whole output (expert.pyx:862-870);
    # with current SciPy a NaN query already raises inside cKDTree.query, for the reference as for us
    I_bad = I[:10].copy()
    I_bad[3] = n
    o, _ = s.interpolate(xq[:10], diff=0, I=I_bad)
    assert np.isnan(o).all()
    xq2 = xq[:10].copy()
    xq2[3, 0] = np.nan
    with pytest.raises(ValueError):
        s.interpolate(xq2, diff=0)


def test_interpolate_continuous_vs_reference_formula():
    n, k = 1500, 20
    x, hoods, f = parity.make_case(n, 2, k)
    nk, od, kn, w = _uniform_meta(n, k, 3, 0, 2)
    xk, fk = parity.gathered(x, f, hoods)
    fi_g, _, _, s = _run_gpu(2, nk, od, kn, w, x, xk, fk, np.zeros((n, 10)))
    s.prep_interpolate()
    rng = np.random.default_rng(11)
    xq = x[:200] + 1e-3 * rng.uniform(-1, 1, (200, 2))
    r = 0.03
    out, Iout = s.interpolate(xq, mode='continuous', r=r, diff=wlsqm.i2_X)
    from scipy.spatial import cKDTree
    L = cKDTree(xq).query_ball_tree(cKDTree(x), r=r)
    exp = np.empty(200)
    for m in range(200):
        acc = sw = 0.0
        for li in L[m]:
            v = orc.interpolate_fit(x[li], fi_g[li], 2, 3, xq[m:m + 1], wlsqm.i2_X)[0]
            d2 = ((xq[m] - x[li]) ** 2).sum()
            ww = (1 - np.sqrt(d2 / r ** 2)) ** 2
            acc += ww * v
            sw += ww
        exp[m] = acc / sw
    assert np.allclose(out, exp, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_interpolate_modes_vs_reference_golden(dim):
    """ExpertSolver.interpolate, mode='continuous' and mode='nearest', against the outputs of the UNMODIFIED reference
    (tests/golden/golden_continuous.npz, expert.pyx:830-985): (a) with the reference's coefficients adopted by an
    all-known solver, which isolates search + evaluation + weighting (1e-12 of the largest output); (b) end to end
    through our own fit (the fit's noise floor on top).  Queries with no model within r give NaN, like the reference."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    import make_golden_continuous as mg
    c = mg.inputs(dim)
    g = parity.golden_continuous(dim)
    assert g is not None, "tests/golden/golden_continuous.npz is missing"
    n, no = c["n"], c["no"]
    nk, od, kn, wm = c["meta"]
    # (a) the reference's coefficients, adopted through a solve with every DOF known
    s = wlsqm.ExpertSolver(dim, nk, od, np.full(n, (1 << no) - 1, np.int64), wm)
    s.prepare(c["x"], c["xk"])
    fi = g["fi_ref"].copy()
    s.solve(c["fk"], fi)
    assert np.array_equal(fi, g["fi_ref"])
    s.prep_interpolate()
    # (b) our own fit
    s_own = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    s_own.prepare(c["x"], c["xk"])
    fi_own = c["fi0"].copy()
    s_own.solve(c["fk"], fi_own)
    s_own.prep_interpolate(search='gpu')
    for d in c["diffs"]:
        refc, refn = g["continuous_diff%d" % d], g["nearest_diff%d" % d]
        outc, Ic = s.interpolate(c["xq"], mode='continuous', r=c["r"], diff=d)
        assert Ic.shape == () and Ic.item() is None
        assert np.array_equal(np.isnan(outc), np.isnan(refc)) and int(np.isnan(refc).sum()) == mg.NFAR
        ok = ~np.isnan(refc)
        scale = np.abs(refc[ok]).max()
        assert np.abs(outc[ok] - refc[ok]).max() <= 1e-12 * scale, (dim, d, np.abs(outc[ok] - refc[ok]).max() / scale)
        outn, In = s.interpolate(c["xq"], mode='nearest', diff=d)
        assert np.array_equal(In, g["I_nearest"]) and In.dtype == np.int_
        assert np.abs(outn - refn).max() <= 1e-12 * np.abs(refn).max(), (dim, d)
        # end to end: the fit's own noise (order <= 3 here; 1D has degenerate spacing) on top
        oc2, _ = s_own.interpolate(c["xq"], mode='continuous', r=c["r"], diff=d)
        on2, In2 = s_own.interpolate(c["xq"], mode='nearest', diff=d)
        assert np.array_equal(In2, g["I_nearest"])           # device-side nearest-model search == cKDTree
        tol = (1e-9 if d == 0 else 1e-6) * (100.0 if dim == 1 else 1.0)
        assert np.array_equal(np.isnan(oc2), np.isnan(refc))
        assert np.abs(oc2[ok] - refc[ok]).max() <= tol * scale, (dim, d, np.abs(oc2[ok] - refc[ok]).max() / scale)
        # (the far-away queries extrapolate the nearest model over 1000 spacings, which amplifies the fit's noise: they
        # are compared with the reference's coefficients above, not here)
        assert np.abs(on2[ok] - refn[ok]).max() <= tol * np.abs(refn[ok]).max(), (dim, d)


def test_simple_api_many_equals_expert_and_single():
    """fit_2D_many_parallel == ExpertSolver == loop of fit_2D (tests/test_simple.py:132-168, test_parallel.py:35-66)"""
    n, k = 64, 12
    x, hoods, f = parity.make_case(n, 2, k, unit_box=True)
    nk, od, kn, w = _uniform_meta(n, k, 2, wlsqm.b2_F, wlsqm.WEIGHT_CENTER)
    xk, fk = parity.gathered(x, f, hoods)
    fi0 = np.zeros((n, 6))
    fi0[:, 0] = f
    fi_many = fi0.copy()
    sens = np.zeros((n, k, 6))
    assert wlsqm.fit_2D_many_parallel(xk, fk, nk, x, fi_many, sens, 1, od, kn, w, ntasks=4) == 0
    fi_e, sens_e, _, _ = _run_gpu(2, nk, od, kn, w, x, xk, fk, fi0, do_sens=True)
    assert np.array_equal(fi_many, fi_e)
    assert np.array_equal(np.nan_to_num(sens), np.nan_to_num(sens_e))
    for j in range(0, n, 9):
        fi1 = fi0[j].copy()
        s1 = np.zeros((k, 6))
        wlsqm.fit_2D(xk[j], fk[j], x[j], fi1, s1, do_sens=1, order=2, knowns=wlsqm.b2_F,
                     weighting_method=wlsqm.WEIGHT_CENTER)
        assert np.array_equal(fi1, fi_many[j])
    fi_it = fi0.copy()
    it = wlsqm.fit_2D_iterative_many_parallel(xk, fk, nk, x, fi_it, None, 0, od, kn, w, max_iter=5)
    fi_o, _, it_o, _ = parity.oracle_solve(2, nk, od, kn, w, x, xk, fk, fi0, 2, False, 5)
    assert it == it_o
    assert np.abs(fi_it - fi_o).max() < 1e-9


def test_mgeneral_drivers_vs_numpy():
    from wlsqm_b200.utils import lapackdrivers as ld
    rng = np.random.default_rng(0)
    for n in (1, 3, 15, 36, 70):
        nlhs = 300
        A = np.asfortranarray(rng.standard_normal((n, n, nlhs)))
        b = np.asfortranarray(rng.standard_normal((n, nlhs)))
        x_ref = np.stack([np.linalg.solve(A[:, :, l], b[:, l]) for l in range(nlhs)], axis=1)
        LU = A.copy(order='F')
        ipiv = np.zeros((n, nlhs), dtype=np.int32, order='F')
        ld.mgeneralfactor(LU, ipiv)
        assert ipiv.min() >= 1 and ipiv.max() <= n
        # the factorisation matches the oracle's dgetf2 restatement pivot for pivot
        LUo = A.copy(order='F')
        ipo = np.zeros_like(ipiv)
        orc.mgetrf(LUo, ipo)
        assert np.array_equal(ipiv, ipo)
        assert np.allclose(LU, LUo, rtol=1e-9, atol=1e-11)
        x = b.copy(order='F')
        ld.mgeneralfactoredp(LU, ipiv, x, 4)
        assert np.allclose(x, x_ref, rtol=1e-8, atol=1e-9)
        A2, b2 = A.copy(order='F'), b.copy(order='F')
        ld.mgeneralp(A2, b2, 8)
        assert np.array_equal(b2, x)


def test_msymmetric_drivers_vs_reference_golden_and_lapack():
    """batched Bunch-Kaufman drivers: outputs of the unmodified reference (tests/golden/golden_sym.npz),
    the oracle's dsytf2 / dsytrs restatement, and numpy on larger batches; host arrays and CUDA tensors"""
    from pathlib import Path
    from wlsqm_b200.utils import lapackdrivers as ld
    g = np.load(Path(__file__).resolve().parent / "golden" / "golden_sym.npz")
    for nm in (str(s) for s in g["names"]):
        A, b = g[f"{nm}/A"], g[f"{nm}/b"]
        n, _, nlhs = A.shape
        F = np.asfortranarray(A.copy())
        low_before = np.tril(F.transpose(2, 0, 1), -1).copy()
        ipiv = np.zeros((n, nlhs), dtype=np.int32, order="F")
        ld.msymmetricfactor(F, ipiv)
        assert np.array_equal(ipiv, g[f"{nm}/ipiv"]), nm                       # pivot sequence: exact
        iu = np.triu_indices(n)
        for l in range(nlhs):
            assert np.allclose(F[:, :, l][iu], g[f"{nm}/F"][:, :, l][iu], rtol=1e-10, atol=1e-12), (nm, l)
        assert np.array_equal(np.tril(F.transpose(2, 0, 1), -1), low_before)    # strict lower triangle not referenced
        x = np.asfortranarray(b.copy())
        ld.msymmetricfactoredp(F, ipiv, x, 4)
        sc = np.abs(g[f"{nm}/x"]).max(axis=0)
        assert (np.abs(x - g[f"{nm}/x"]).max(axis=0) <= 1e-9 * sc).all(), nm
        A2, x2 = np.asfortranarray(A.copy()), np.asfortranarray(b.copy())
        ld.msymmetricp(A2, x2, 8)
        assert np.array_equal(x2, x)
        S = np.asfortranarray(g[f"{nm}/G"].copy())
        ld.msymmetrizep(S, 2)
        assert np.array_equal(S, g[f"{nm}/S"]), nm
    # larger batches against numpy and the restatement (pivots exact), incl. sizes beyond one warp's lane count
    rng = np.random.default_rng(1)
    torch = pytest.importorskip("torch")
    for n in (1, 2, 5, 15, 36, 70):
        nlhs = 257
        A = rng.standard_normal((n, n, nlhs))
        A = np.asfortranarray(0.5 * (A + A.transpose(1, 0, 2)))
        A[np.arange(n), np.arange(n), ::3] *= 1e-3                              # provoke 2x2 blocks
        b = np.asfortranarray(rng.standard_normal((n, nlhs)))
        x_ref = np.stack([np.linalg.solve(A[:, :, l], b[:, l]) for l in range(nlhs)], axis=1)
        F, ipiv = A.copy(order="F"), np.zeros((n, nlhs), dtype=np.int32, order="F")
        ld.msymmetricfactor(F, ipiv)
        Fo, ipo = A.copy(order="F"), np.zeros_like(ipiv)
        orc.msytrf(Fo, ipo)
        assert np.array_equal(ipiv, ipo), n
        x = b.copy(order="F")
        ld.msymmetricfactored(F, ipiv, x)
        err = np.abs(x - x_ref).max(axis=0) / np.abs(x_ref).max(axis=0)
        assert np.median(err) < 1e-12 and err.max() < 1e-7, (n, err.max())
        # CUDA tensors, zero copy (Fortran order = transposed view of a C-contiguous tensor)
        A_t = torch.from_numpy(np.ascontiguousarray(A.transpose(2, 1, 0))).cuda().permute(2, 1, 0)
        b_t = torch.from_numpy(np.ascontiguousarray(b.T)).cuda().t()
        ld.msymmetric(A_t, b_t)
        torch.cuda.synchronize()
        assert np.array_equal(b_t.cpu().numpy(), x)


def test_batched_drivers_every_size():
    """mgeneral* / msymmetric* for EVERY matrix size from 1 to 72 (the kernels switch layouts at 16, 32 and 64 rows, and the
    batch count is not a multiple of anything): pivots exact against the oracle's dgetf2 / dsytf2 restatements, solutions
    against numpy"""
    from wlsqm_b200.utils import lapackdrivers as ld
    rng = np.random.default_rng(5)
    for n in range(1, 73):
        nlhs = 37
        A = np.asfortranarray(rng.standard_normal((n, n, nlhs)))
        b = np.asfortranarray(rng.standard_normal((n, nlhs)))
        x_ref = np.stack([np.linalg.solve(A[:, :, l], b[:, l]) for l in range(nlhs)], axis=1)
        LU, ipiv = A.copy(order="F"), np.zeros((n, nlhs), dtype=np.int32, order="F")
        ld.mgeneralfactor(LU, ipiv)
        LUo, ipo = A.copy(order="F"), np.zeros_like(ipiv)
        orc.mgetrf(LUo, ipo)
        assert np.array_equal(ipiv, ipo), n
        assert np.allclose(LU, LUo, rtol=1e-8, atol=1e-10), n
        x = b.copy(order="F")
        ld.mgeneralfactored(LU, ipiv, x)
        err = np.abs(x - x_ref).max(axis=0) / np.abs(x_ref).max(axis=0)
        assert np.median(err) < 1e-11 and err.max() < 1e-6, (n, err.max())
        S = np.asfortranarray(0.5 * (A + A.transpose(1, 0, 2)))
        S[np.arange(n), np.arange(n), ::3] *= 1e-3
        xs_ref = np.stack([np.linalg.solve(S[:, :, l], b[:, l]) for l in range(nlhs)], axis=1)
        F, ips = S.copy(order="F"), np.zeros((n, nlhs), dtype=np.int32, order="F")
        ld.msymmetricfactor(F, ips)
        Fo, ipso = S.copy(order="F"), np.zeros_like(ips)
        orc.msytrf(Fo, ipso)
        assert np.array_equal(ips, ipso), n
        xs = b.copy(order="F")
        ld.msymmetricfactored(F, ips, xs)
        err = np.abs(xs - xs_ref).max(axis=0) / np.abs(xs_ref).max(axis=0)
        assert np.median(err) < 1e-11 and err.max() < 1e-6, (n, err.max())


def test_batched_scalers_vs_reference_golden():
    """mdo_rescale / do_rescale / rescale_* (SURVEY 8f item 4): one launch over a batch of matrices against the unmodified
    reference's do_rescale applied matrix by matrix (tests/golden/golden_scale.npz) -- scale vectors and scaled matrices
    bit for bit, for all six algorithms, square / rectangular / Gram-like batches; host arrays and CUDA tensors; the
    DGEEQU failure (zero row) raises LinAlgError like the reference and leaves that matrix alone."""
    from wlsqm_b200.utils import lapackdrivers as ld
    torch = pytest.importorskip("torch")
    g = np.load(parity.GOLDEN_DIR / "golden_scale.npz")
    for nm in (str(s) for s in g["names"]):
        A = g[f"{nm}/A"]
        for algo in ld.ScalingAlgo:
            M = np.asfortranarray(A.copy())
            rs, cs = ld.mdo_rescale(M, algo)
            a = int(algo)
            assert np.array_equal(rs, g[f"{nm}/rs{a}"]), (nm, a)
            assert np.array_equal(cs, g[f"{nm}/cs{a}"]), (nm, a)
            assert np.array_equal(M, g[f"{nm}/S{a}"]), (nm, a)
            # CUDA tensors, zero copy (Fortran order = permuted view of a C-contiguous tensor)
            M_t = torch.from_numpy(np.ascontiguousarray(A.transpose(2, 1, 0))).cuda().permute(2, 1, 0)
            rs_t, cs_t = ld.mdo_rescalep(M_t, a, ntasks=4)
            torch.cuda.synchronize()
            assert np.array_equal(rs_t.cpu().numpy(), rs) and np.array_equal(cs_t.cpu().numpy(), cs)
            assert np.array_equal(M_t.cpu().numpy(), M)
        # the reference's single-matrix entry points
        for fn, a in ((ld.rescale_columns, 1), (ld.rescale_rows, 2), (ld.rescale_twopass, 3), (ld.rescale_ruiz2001, 4),
                      (ld.rescale_scalgm, 5), (ld.rescale_dgeequ, 6)):
            M1 = np.asfortranarray(A[:, :, 1].copy())
            r1, c1 = fn(M1)
            assert np.array_equal(r1, g[f"{nm}/rs{a}"][:, 1]) and np.array_equal(c1, g[f"{nm}/cs{a}"][:, 1])
            assert np.array_equal(M1, g[f"{nm}/S{a}"][:, :, 1])
    # Ruiz keeps a symmetric matrix symmetric (the reference's tests/test_lapackdrivers.py:90-98)
    S = np.asfortranarray(g["gram15/A"][:, :, 3].copy())
    ld.do_rescale(S, ld.ScalingAlgo.ALGO_RUIZ2001)
    assert np.allclose(S, S.T, rtol=0, atol=1e-15)
    # DGEEQU on a matrix with a zero row: LinAlgError (tests/test_lapackdrivers.py:109-115); the other matrices are scaled
    B = np.asfortranarray(g["sq3/A"][:, :, :4].copy())
    B[1, :, 2] = 0.0
    B0 = B.copy()
    with pytest.raises(np.linalg.LinAlgError):
        ld.mdo_rescale(B, ld.ScalingAlgo.ALGO_DGEEQU)
    assert np.array_equal(B[:, :, 2], B0[:, :, 2]) and np.array_equal(B[:, :, 0], g["sq3/S6"][:, :, 0])
    with pytest.raises(ValueError, match="Unknown algorithm"):
        ld.do_rescale(np.asfortranarray(np.eye(3)), 9)


def test_single_system_drivers_vs_lapack():
    """the reference's single-system entry points (general, symmetric, generalfactor/-factored, symmetricfactor/-factored,
    tridiag; lapackdrivers.pyx:854-1050, 1395-1462) on top of the batched kernels; tridiag = DGTSV incl. its row
    interchanges and its in-place factors, against SciPy's LAPACK"""
    from wlsqm_b200.utils import lapackdrivers as ld
    from scipy.linalg import lapack
    rng = np.random.default_rng(5)
    n = 7
    A = rng.standard_normal((n, n)); b = rng.standard_normal(n)
    x_ref = np.linalg.solve(A, b)
    A1, b1 = np.asfortranarray(A.copy()), b.copy()
    ld.general(A1, b1)
    assert np.allclose(b1, x_ref, atol=1e-12)
    LU = np.asfortranarray(A.copy())
    ipiv = ld.generalfactor(LU)
    lu_ref, piv_ref, _ = lapack.dgetrf(A)
    assert np.array_equal(ipiv, piv_ref + 1) and np.allclose(LU, lu_ref, atol=1e-13)
    b2 = b.copy()
    ld.generalfactored(LU, ipiv, b2)
    assert np.allclose(b2, x_ref, atol=1e-12)
    S = 0.5 * (A + A.T) + n * np.eye(n)
    S1, b3 = np.asfortranarray(S.copy()), b.copy()
    ld.symmetric(S1, b3)
    assert np.allclose(b3, np.linalg.solve(S, b), atol=1e-12)
    S2 = np.asfortranarray(S.copy())
    ip2 = ld.symmetricfactor(S2)
    b4 = b.copy()
    ld.symmetricfactored(S2, ip2, b4)
    assert np.allclose(b4, np.linalg.solve(S, b), atol=1e-12)
    # the reference's own example (tests/test_lapackdrivers.py:26-39): a[0] IS DL[0]
    a = np.array([0.0, -1.0, -1.0, -1.0]); d = np.full(4, 2.0); c = np.array([-1.0, -1.0, -1.0, 0.0]); x = np.array([1.0, 0.0, 0.0, 1.0])
    assert ld.tridiag(a, d, c, x) == 0
    assert np.allclose(x, [0.625, 0.25, 0.5, 0.75], atol=1e-14)
    # random systems that need row interchanges: solution AND the overwritten factors equal DGTSV's
    for trial in range(5):
        m = 9
        dl, dd, du, rhs = rng.standard_normal(m), 0.1 * rng.standard_normal(m), rng.standard_normal(m), rng.standard_normal(m)
        dl2, d2, du2, x2, info = lapack.dgtsv(dl[:m - 1].copy(), dd.copy(), du[:m - 1].copy(), rhs.copy())
        assert info == 0
        a_, b_, c_, x_ = dl.copy(), dd.copy(), du.copy(), rhs.copy()
        ld.tridiag(a_, b_, c_, x_)
        assert np.allclose(x_, x2, rtol=1e-12, atol=1e-13)
        assert np.allclose(b_, d2, rtol=1e-13) and np.allclose(a_[:m - 1], dl2, rtol=1e-13) and np.allclose(c_[:m - 1], du2, rtol=1e-13)
