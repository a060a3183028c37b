"""Edge cases of the hot path, checked against the unmodified reference's behaviour (probed in the build container,
restated here as assertions; the oracle covers the arithmetic): empty and one-case batches, neighbourhoods without
points, underdetermined fits next to healthy ones, large neighbourhoods, pitched / strided inputs."""
import numpy as np
import pytest

import parity
import oracle as orc

pytestmark = pytest.mark.gpu

wlsqm = pytest.importorskip("wlsqm_b200")


def _meta(n, k, order, knowns, wm):
    return (np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, knowns, np.int64), np.full(n, wm, np.int32))


def test_empty_batch_is_refused_like_the_reference():
    e32, e64 = np.zeros(0, np.int32), np.zeros(0, np.int64)
    with pytest.raises(ValueError, match="max_cases > 0"):          # infra.pyx CaseManager_new
        wlsqm.ExpertSolver(2, e32, e32, e64, e32)
    for fn in (wlsqm.fit_2D_many, wlsqm.fit_2D_many_parallel, wlsqm.fit_2D_iterative_many):
        with pytest.raises(ValueError, match="max_cases > 0"):
            fn(np.zeros((0, 5, 2)), np.zeros((0, 5)), e32, np.zeros((0, 2)), np.zeros((0, 6)), None, 0, e32, e64, e32)


def test_single_case_and_single_neighbour():
    # order 0 with one neighbour: the fit is that neighbour's value (reference: fit_2D(...) -> fi = [3.])
    fi = np.zeros(1)
    assert wlsqm.fit_2D(np.array([[0.1, 0.2]]), np.array([3.0]), np.array([0.0, 0.0]), fi, None, 0, 0, 0, 1) == 0
    assert abs(fi[0] - 3.0) <= 2 * np.finfo(float).eps * 3.0      # (the equilibration rounds 1 / sqrt(w))
    # a batch of one case through every entry point agrees with the oracle
    x, hoods, f = parity.make_case(50, 2, 12)
    xk, fk = parity.gathered(x, f, hoods)
    m = _meta(1, 12, 2, 1, 2)
    fi1 = np.zeros((1, 6)); fi1[0, 0] = f[0]
    fo = fi1.copy()
    wlsqm.fit_2D_many(xk[:1], fk[:1], m[0], x[:1], fi1, None, 0, m[1], m[2], m[3])
    orc.fit_many(2, xk[:1], fk[:1], m[0], x[:1], fo, None, False, m[1], m[2], m[3])
    assert np.allclose(fi1, fo, rtol=1e-10, atol=1e-12)


def test_cases_without_neighbours_give_nan_and_leave_the_rest_alone():
    """nk = 0: the reference's normal matrix is all zeros and every unknown comes out NaN; other cases are unaffected"""
    n, k = 64, 12
    x, hoods, f = parity.make_case(n, 2, k)
    xk, fk = parity.gathered(x, f, hoods)
    nk, od, kn, wm = _meta(n, k, 2, 0, 1)
    nk[::8] = 0
    fi = np.zeros((n, 6))
    s = wlsqm.ExpertSolver(2, nk, od, kn, wm)
    s.prepare(x, xk)
    s.solve(fk, fi)
    assert np.isnan(fi[::8]).all()
    good = np.ones(n, bool); good[::8] = False
    so = orc.OracleSolver(2, nk[good], od[good], kn[good], wm[good])
    so.prepare(x[good], xk[good])
    fo = np.zeros((good.sum(), 6))
    so.solve(fk[good], fo)
    assert np.isfinite(fi[good]).all()
    assert np.abs(fi[good] - fo).max() <= 1e-9 * np.abs(fo).max()


def test_underdetermined_fits_do_not_disturb_their_neighbours_in_the_batch():
    """nk < number of unknowns: the matrix is singular, the reference returns whatever LU of round-off gives (finite
    garbage or inf/NaN, no error).  Required here: no error, and the healthy cases of the same batch are untouched."""
    n, k = 200, 30
    x, hoods, f = parity.make_case(n, 2, k)
    xk, fk = parity.gathered(x, f, hoods)
    nk, od, kn, wm = _meta(n, k, 4, 0, 2)
    nk[5::10] = 7                       # 7 neighbours for 15 unknowns
    fi = np.zeros((n, 15))
    s = wlsqm.ExpertSolver(2, nk, od, kn, wm)
    s.prepare(x, xk)
    s.solve(fk, fi)
    good = np.ones(n, bool); good[5::10] = False
    so = orc.OracleSolver(2, nk[good], od[good], kn[good], wm[good])
    so.prepare(x[good], xk[good])
    fo = np.zeros((good.sum(), 15))
    so.solve(fk[good], fo)
    rep = parity.wl.parity_report(fi[good], fo, 2, 4)
    assert rep[0][2] < 1e-9 and rep[1][1] < 1e-9, rep


@pytest.mark.parametrize("dim,order,k", [(2, 4, 200), (3, 4, 150), (1, 4, 300), (2, 2, 500)])
def test_large_neighbourhoods(dim, order, k):
    """neighbourhoods far larger than the usual 25-60 points (several 32-column blocks per fit, operator blocks of
    tens of KB): same answers as the oracle"""
    n = 120
    x, hoods, f = parity.make_case(max(n, k + 50), dim, k)
    x, hoods, f = x, hoods[:n], f
    xi = x[:n]
    xk, fk = np.ascontiguousarray(x[hoods]), np.ascontiguousarray(f[hoods])
    no = wlsqm.number_of_dofs(dim, order)
    nk, od, kn, wm = _meta(n, k, order, 1, 2)
    fi0 = np.zeros((n, no)); fi0[:, 0] = f[:n]
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, do_sens=True)
    s.prepare(xi, xk)
    fi = fi0.copy()
    sens = np.zeros((n, k, no))
    s.solve(fk, fi, sens)
    ref, sens_o, _, _ = parity.oracle_solve(dim, nk, od, kn, wm, xi, xk, fk, fi0, do_sens=True)
    a, b = parity.permuted_self_noise(dim, nk, od, kn, wm, xi, xk, fk, fi0)
    print(parity.check_against_floor(fi, ref, b + (ref - a), dim, order, f"k={k}"))
    parity.check_sens(sens, sens_o, f"k={k}")


def test_pitched_and_strided_host_inputs():
    """the memoryview layouts of simple.pyx:149-159: leading axes arbitrarily strided, fk strided on both axes"""
    n, k = 300, 24
    x, hoods, f = parity.make_case(n, 2, k)
    xk, fk = parity.gathered(x, f, hoods)
    no = 10
    nk, od, kn, wm = _meta(n, k, 3, 1, 1)
    fi_ref = np.zeros((n, no)); fi_ref[:, 0] = f
    wlsqm.fit_2D_many(xk, fk, nk, x, fi_ref, None, 0, od, kn, wm)
    # every array embedded in a larger one
    xk_big = np.zeros((2 * n, k + 3, 2)); xk_big[::2, :k] = xk
    fk_big = np.zeros((n, 2 * k + 1)); fk_big[:, :2 * k:2] = fk
    x_big = np.zeros((n, 5)); x_big[:, :2] = x
    fi_big = np.full((n, no + 4), 7.0); fi_big[:, 0] = f
    wlsqm.fit_2D_many(xk_big[::2, :k], fk_big[:, :2 * k:2], nk, x_big[:, :2], fi_big[:, :no], None, 0, od, kn, wm)
    assert np.array_equal(fi_big[:, :no], fi_ref)
    assert (fi_big[:, no:] == 7.0).all()          # columns beyond the model are not touched


def test_large_pageable_arrays_take_the_threaded_staging_path():
    """ordinary numpy arrays above 64 MB go through the page-locked rings with threaded host copies
    (csrc/wlsqm_host.cu): same bits as page-locked arrays and as CUDA tensors, pitched fi rows left intact"""
    torch = pytest.importorskip("torch")
    n, k, no = 320_000, 30, 15            # fk = 76.8 MB
    g = torch.Generator(device="cuda").manual_seed(3)
    xi_d = 20.0 * torch.rand((n, 2), dtype=torch.float64, device="cuda", generator=g)
    xk_d = xi_d[:, None, :] + 0.015 * (2 * torch.rand((n, k, 2), dtype=torch.float64, device="cuda", generator=g) - 1)
    fk_d = torch.sin(xk_d[..., 0]) * torch.cos(xk_d[..., 1])
    nk, od, kn, wm = _meta(n, k, 4, 0, 1)
    s = wlsqm.ExpertSolver(2, nk, od, kn, wm)
    xk_h = xk_d.cpu().numpy()             # 153.6 MB pageable: prepare stages it through the rings too
    s.prepare(xi_d.cpu().numpy(), xk_h)
    fi_d = torch.zeros((n, no), dtype=torch.float64, device="cuda")
    s.solve(fk_d, fi_d)
    ref = fi_d.cpu().numpy()
    fk_h = fk_d.cpu().numpy()
    fi_big = np.full((n, no + 3), 5.0)    # pitched rows: columns beyond the model must survive
    s.solve(fk_h, fi_big[:, :no])
    assert np.array_equal(fi_big[:, :no], ref)
    assert (fi_big[:, no:] == 5.0).all()
    fk_p, fi_p = wlsqm.pinned_empty((n, k)), wlsqm.pinned_empty((n, no))
    fk_p[...] = fk_h
    fi_p[...] = 0.0
    s.solve(fk_p, fi_p)
    assert np.array_equal(fi_p, ref)
    # interpolate with large pageable query arrays (x 144 MB, I 72 MB, out 72 MB) == the device-tensor path
    nq = 9_000_000
    I_d = torch.randint(0, n, (nq,), device="cuda", generator=g)
    xq_d = xi_d[I_d] + 0.004 * (2 * torch.rand((nq, 2), dtype=torch.float64, device="cuda", generator=g) - 1)
    s.tree = object()
    out_d, _ = s.interpolate(xq_d, diff=wlsqm.i2_XY, I=I_d)
    out_h, I_back = s.interpolate(xq_d.cpu().numpy(), diff=wlsqm.i2_XY, I=I_d.cpu().numpy())
    assert isinstance(out_h, np.ndarray) and np.array_equal(out_h, out_d.cpu().numpy())
    del I_d, xq_d, out_d, out_h
    # the same prepared state from device arrays gives the same operators
    s2 = wlsqm.ExpertSolver(2, nk, od, kn, wm)
    s2.prepare(xi_d, xk_d)
    fi2 = torch.zeros_like(fi_d)
    s2.solve(fk_d, fi2)
    assert torch.equal(fi2, fi_d)


# ---- robustness of the extension entry points (hood index lists, search grid, CUDA streams) --------------------
def test_ragged_hoods_padding_is_never_dereferenced_and_bad_indices_raise():
    """hoods rows with nk[i] < max nk carry padding (cKDTree and the device search both report a missing neighbour as
    the index n): only the first nk[i] entries are used; a USED index outside [0, npoints) raises ValueError where the
    reference's caller-side gather x[hoods] raises IndexError."""
    n, k, dim, order = 600, 14, 2, 2
    x, hoods, f = parity.make_case(n, dim, k)
    rng = np.random.default_rng(4)
    nk = rng.integers(8, k + 1, n).astype(np.int32)
    nk[0] = k
    od, kn, wm = np.full(n, order, np.int32), np.full(n, 1, np.int64), np.full(n, 2, np.int32)
    ragged = hoods.copy()
    pad = np.array([n, -1, 2 ** 31 - 1, n + 12345], np.int32)
    for i in range(n):
        ragged[i, nk[i]:] = pad[rng.integers(0, 4, k - nk[i])]
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    s.prepare_hoods(x, ragged)
    fi = np.zeros((n, 6)); fi[:, 0] = f
    s.solve_hoods(f, fi)
    # == the pre-gathered path on the same neighbourhoods (padding slots hold anything there, too)
    xk, fk = parity.gathered(x, f, hoods)
    s2 = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    s2.prepare(x, xk)
    fi2 = np.zeros((n, 6)); fi2[:, 0] = f
    s2.solve(fk, fi2)
    assert np.array_equal(fi, fi2)
    bad = hoods.copy()
    bad[17, 3] = n                      # a used slot
    s3 = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    with pytest.raises(ValueError, match="hoods"):
        s3.prepare_hoods(x, bad)
    assert not s3.ready
    bad[17, 3] = -2
    with pytest.raises(ValueError, match="hoods"):
        s3.prepare_hoods(x, bad)
    s3.prepare_hoods(x, hoods)          # and the solver is still usable afterwards
    torch = pytest.importorskip("torch")
    g = wlsqm.gather(torch.from_numpy(f).cuda(), torch.from_numpy(bad).cuda())
    torch.cuda.synchronize()
    gh = g.cpu().numpy()
    assert np.isnan(gh[17, 3]) and np.isfinite(np.delete(gh.ravel(), 17 * k + 3)).all()


def test_search_grid_with_non_finite_coordinates_does_not_hang():
    """+/-inf coordinates are treated like NaN ones (nobody's neighbour, they do not shape the cell grid); a cloud with
    no finite point, or an extent whose volume overflows, is a ValueError -- not an endless loop (run with a timeout)."""
    x, hoods, f = parity.make_case(2000, 2, 8)
    xb = x.copy()
    xb[5, 0] = np.inf
    xb[9, 1] = -np.inf
    xb[11, 0] = np.nan
    h = wlsqm.knn_hoods(xb, 8)
    good = np.ones(2000, bool); good[[5, 9, 11]] = False
    assert not np.isin(h[good], [5, 9, 11]).any()
    from scipy.spatial import cKDTree
    idx = np.nonzero(good)[0]
    ref = idx[cKDTree(x[good]).query(x[good], 9)[1][:, 1:]]
    assert np.array_equal(h[good], ref)
    with pytest.raises(ValueError):
        wlsqm.PointGrid(np.full((10, 2), np.inf))
    huge = np.array([[-1e200, -1e200, -1e200], [1e200, 1e200, 1e200], [0.0, 0.0, 0.0]])
    with pytest.raises(ValueError):
        wlsqm.PointGrid(huge)


def test_work_is_ordered_against_torch_side_streams():
    """inputs produced on a NON-default torch stream right before the call, outputs consumed right after it on that
    stream: one-shot fits, the search grid and a solver that changes streams between prepare and solve"""
    torch = pytest.importorskip("torch")
    n, k = 200_000, 12
    x, hoods, f = parity.make_case(n, 2, k, unit_box=True)
    nk, od, kn, wm = _meta(n, k, 2, 1, 2)
    xk, fk = parity.gathered(x, f, hoods)
    fi0 = np.zeros((n, 6)); fi0[:, 0] = f
    ref = fi0.copy()
    wlsqm.fit_2D_many_parallel(xk, fk, nk, x, ref, None, 0, od, kn, wm)
    # (the one-shot kernel and the prepare + solve path round differently: each has its own host-array result)
    ref_expert = fi0.copy()
    s0 = wlsqm.ExpertSolver(2, nk, od, kn, wm)
    s0.prepare(x, xk)
    s0.solve(fk, ref_expert)
    del s0
    xk_h, fk_h, x_h, fi_h = (torch.from_numpy(a).pin_memory() for a in (xk, fk, x, fi0))
    side, side2 = torch.cuda.Stream(), torch.cuda.Stream()
    for rep in range(3):
        with torch.cuda.stream(side):
            # a long copy chain queued on the side stream; the library call must wait for it
            junk = torch.empty(64 << 20, dtype=torch.float64, device="cuda")
            for _ in range(4):
                junk.mul_(1.0000001)
            xk_d, fk_d, x_d = (t.to("cuda", non_blocking=True) for t in (xk_h, fk_h, x_h))
            fi_d = fi_h.to("cuda", non_blocking=True)
            wlsqm.fit_2D_many_parallel(xk_d, fk_d, nk, x_d, fi_d, None, 0, od, kn, wm)
            out = fi_d.cpu().numpy()
        assert np.array_equal(out, ref)
        with torch.cuda.stream(side):
            junk.mul_(1.0000001)
            x_d2 = x_h.to("cuda", non_blocking=True)
            hd = wlsqm.knn_hoods(x_d2, k)
        assert np.array_equal(hd.cpu().numpy(), hoods)
        # prepare on one stream, solve on another
        s = wlsqm.ExpertSolver(2, nk, od, kn, wm)
        with torch.cuda.stream(side):
            junk.mul_(1.0000001)
            xk_d, x_d = xk_h.to("cuda", non_blocking=True), x_h.to("cuda", non_blocking=True)
            s.prepare(x_d, xk_d)
        with torch.cuda.stream(side2):
            side2.wait_stream(side)          # (the tensors were produced on `side`)
            fk_d, fi_d = fk_h.to("cuda", non_blocking=True), fi_h.to("cuda", non_blocking=True)
            s.solve(fk_d, fi_d)
            out = fi_d.cpu().numpy()
        assert np.array_equal(out, ref_expert)
        del junk


def test_host_xk_with_a_longer_last_axis_and_list_outputs():
    """the reference's double[:,:,::contiguous] xk accepts an (n, k, 3) array for a 2D fit and reads the first two
    columns; a list passed where results are written is refused (the reference's memoryviews raise, too)"""
    n, k = 300, 12
    x, hoods, f = parity.make_case(n, 2, k, unit_box=True)
    nk, od, kn, wm = _meta(n, k, 2, 1, 2)
    xk, fk = parity.gathered(x, f, hoods)
    xk3 = np.concatenate([xk, np.full((n, k, 1), 7.0)], axis=2)
    fi_a = np.zeros((n, 6)); fi_a[:, 0] = f
    fi_b = fi_a.copy(); fi_c = fi_a.copy()
    wlsqm.fit_2D_many_parallel(xk, fk, nk, x, fi_a, None, 0, od, kn, wm)
    wlsqm.fit_2D_many_parallel(xk3, fk, nk, x, fi_b, None, 0, od, kn, wm)
    assert np.array_equal(fi_a, fi_b)
    s = wlsqm.ExpertSolver(2, nk, od, kn, wm)
    s.prepare(x, xk3)
    s.solve(fk, fi_c)
    fi_d = fi_a.copy()
    fi_d[:, 1:] = 0.0
    s.prepare(x, xk)                     # (prepare + solve rounds differently from the one-shot kernel: like with like)
    s.solve(fk, fi_d)
    assert np.array_equal(fi_c, fi_d)
    with pytest.raises((TypeError, ValueError)):
        s.solve(fk, fi_c.tolist())


@pytest.mark.parametrize("hetero", [False, True])
def test_host_sens_and_fi_go_back_in_chunks(hetero, monkeypatch):
    """host sens / fi of a many-chunk batch: the kernel writes sens into two rotating chunk buffers (no device mirror of the
    whole array) and the chunks go back through the page-locked ring (ordinary numpy memory, per-case nk / no) or by
    direct asynchronous copies (page-locked, dense, uniform); same bits as the device-tensor path"""
    torch = pytest.importorskip("torch")
    monkeypatch.setenv("WLSQM_SOLVE_CHUNK", "4096")
    n, k, dim = 30_000, 14, 2
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    rng = np.random.default_rng(9)
    if hetero:
        od = rng.integers(1, 4, n).astype(np.int32)
        nk = rng.integers(11, k + 1, n).astype(np.int32)
        kn = rng.integers(0, 2, n).astype(np.int64)
        kn[::17] = (1 << 3) - 1          # all DOFs of an order-1 model known where od == 1: a silent no-op case
    else:
        od, nk, kn = np.full(n, 3, np.int32), np.full(n, k, np.int32), np.full(n, 1, np.int64)
    wm = np.full(n, 2, np.int32)
    no = 10
    fi0 = rng.standard_normal((n, no)); fi0[:, 0] = f
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, do_sens=True)
    s.prepare(x, xk)
    fi_d = torch.from_numpy(fi0).cuda()
    sens_d = torch.full((n, k, no), -7.0, dtype=torch.float64, device="cuda")
    s.solve(torch.from_numpy(fk).cuda(), fi_d, sens_d)
    torch.cuda.synchronize()
    ref_fi, ref_sens = fi_d.cpu().numpy(), sens_d.cpu().numpy()
    # ordinary numpy memory
    fi_h, sens_h = fi0.copy(), np.full((n, k, no), -7.0)
    s.solve(fk, fi_h, sens_h)
    assert np.array_equal(fi_h, ref_fi)
    assert np.array_equal(np.isnan(sens_h), np.isnan(ref_sens)) and np.array_equal(np.nan_to_num(sens_h), np.nan_to_num(ref_sens))
    # pitched host sens (rows longer than the model) and page-locked memory
    sens_p = np.full((n, k + 2, no + 3), -7.0)
    fi_h2 = fi0.copy()
    s.solve(fk, fi_h2, sens_p[:, :k, :no])
    assert np.array_equal(np.nan_to_num(sens_p[:, :k, :no]), np.nan_to_num(ref_sens)) and (sens_p[:, k:, :] == -7.0).all() and (sens_p[:, :, no:] == -7.0).all()
    sens_l = wlsqm.pinned_empty((n, k, no)); sens_l[...] = -7.0
    fi_l = wlsqm.pinned_empty((n, no)); fi_l[...] = fi0
    s.solve(fk, fi_l, sens_l)
    assert np.array_equal(fi_l, ref_fi) and np.array_equal(np.nan_to_num(sens_l), np.nan_to_num(ref_sens))
    wlsqm.pinned_free(sens_l); wlsqm.pinned_free(fi_l)


@pytest.mark.parametrize("algo", ["basic", "iterative"])
def test_solution_copy_can_be_dropped_for_device_outputs(algo):
    """keep_solution(False): solve() with a CUDA-tensor fi writes the same bits into the caller's array and nothing into
    the solver's own copy (the reference's Case_set_fi); interpolate() refuses until a solve() that keeps the copy;
    host outputs are unaffected (they are read back from the solver's copy)"""
    torch = pytest.importorskip("torch")
    n, k = 3000, 14
    x, hoods, f = parity.make_case(n, 2, k)
    nk, od, kn, wm = _meta(n, k, 3, 0b10, 2)                      # (a known DOF: fi keeps the caller's value there)
    xk, fk = parity.gathered(x, f, hoods)
    fi0 = np.zeros((n, 10)); fi0[:, 1] = 0.25
    A = wlsqm.ALGO_ITERATIVE if algo == "iterative" else wlsqm.ALGO_BASIC
    s = wlsqm.ExpertSolver(2, nk, od, kn, wm, algorithm=A)
    s.prepare(x, xk)
    want = fi0.copy()
    it_want = s.solve(fk, want)
    xq = x[:50] + 1e-3
    I = np.arange(50, dtype=np.int64)
    s.prep_interpolate()
    out_want, _ = s.interpolate(xq, mode='nearest', diff=0, I=I)

    fk_d = torch.from_numpy(fk).cuda()
    s.keep_solution(False)
    got = torch.from_numpy(fi0).cuda()
    it_got = s.solve(fk_d, got)
    assert np.array_equal(got.cpu().numpy(), want) and it_got == it_want
    with pytest.raises(RuntimeError, match="keeps no copy"):
        s.interpolate(xq, mode='nearest', diff=0, I=I)
    host = fi0.copy()
    s.solve(fk, host)                                             # host output: staged through the copy, which is valid again
    assert np.array_equal(host, want)
    out, _ = s.interpolate(xq, mode='nearest', diff=0, I=I)
    assert np.array_equal(out, out_want)
    s.keep_solution(True)
    got2 = torch.from_numpy(fi0).cuda()
    s.solve(fk_d, got2)
    out, _ = s.interpolate(xq, mode='nearest', diff=0, I=I)
    assert np.array_equal(got2.cpu().numpy(), want) and np.array_equal(out, out_want)


def test_odd_neighbour_counts_in_even_pitched_rows_whose_storage_ends_with_the_last_element():
    """fk rows with an odd element count travel by bulk copy with one element more (the row pitch is even); the last row of a
    caller's strided tensor may have no such element -- same results as from a contiguous copy, knowns included"""
    torch = pytest.importorskip("torch")
    n, k = 4099, 13
    x, hoods, f = parity.make_case(n, 2, k)
    nk, od, kn, wm = _meta(n, k, 3, 0b101, 2)
    xk, fk = parity.gathered(x, f, hoods)
    fi0 = np.zeros((n, 10)); fi0[:, 0] = f; fi0[:, 2] = 0.5
    s = wlsqm.ExpertSolver(2, nk, od, kn, wm)
    s.prepare(torch.from_numpy(x).cuda(), torch.from_numpy(xk).cuda())
    want = torch.from_numpy(fi0).cuda()
    s.solve(torch.from_numpy(fk).cuda(), want)                       # contiguous rows of 13: lane gather (odd pitch)
    base = torch.full(((n - 1) * 14 + 13,), float("nan"), dtype=torch.float64, device="cuda")
    view = base.as_strided((n, k), (14, 1))
    view.copy_(torch.from_numpy(fk).cuda())
    got = torch.from_numpy(fi0).cuda()
    s.solve(view, got)                                               # even pitch: bulk copies, the last row by the lanes
    assert torch.equal(got, want)
    ref, _, _, _ = parity.oracle_solve(2, nk, od, kn, wm, x, xk, fk, fi0)
    assert np.abs(got.cpu().numpy() - ref).max() <= 1e-8 * np.abs(ref).max()
