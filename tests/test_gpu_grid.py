"""GPU tests of the device-side spatial search: neighbourhood construction, nearest-model
search and ball search, validated index for index against scipy.spatial.cKDTree -- the tool the reference's callers
and ExpertSolver use for the same steps (examples/expertsolver_example.py:51-66, expert.pyx:676-681,837,898-911).
Random clouds have distinct distances, so the k nearest neighbours and their order are unique."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

import parity

pytestmark = pytest.mark.gpu

wlsqm = pytest.importorskip("wlsqm_b200")


def _cloud(n, dim, kind, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        x = rng.random((n, dim))
    elif kind == "clustered":        # strongly non-uniform density: many empty cells, some crowded ones
        c = rng.random((20, dim))
        x = c[rng.integers(0, 20, n)] + 0.01 * rng.standard_normal((n, dim))
    else:                            # anisotropic box
        x = rng.random((n, dim)) * np.array([100.0, 1.0, 0.01])[:dim]
    return x[:, 0].copy() if dim == 1 else x


@pytest.mark.parametrize("dim,k,kind,n", [(2, 30, "uniform", 20000), (2, 12, "clustered", 20000), (3, 60, "uniform", 8000),
                                          (3, 20, "anisotropic", 8000), (1, 8, "uniform", 20000), (1, 9, "clustered", 5000),
                                          (2, 1, "uniform", 3000), (2, 100, "uniform", 4000)])
def test_knn_matches_ckdtree(dim, k, kind, n):
    x = _cloud(n, dim, kind)
    x2 = x.reshape(n, -1)
    d_ref, i_ref = cKDTree(x2).query(x2, k + 1)
    g = wlsqm.PointGrid(x)
    d, hoods = g.knn(k, return_distance=True)
    assert hoods.dtype == np.int32 and hoods.shape == (n, k)
    assert np.array_equal(hoods, i_ref[:, 1:])
    assert np.allclose(d, d_ref[:, 1:], rtol=1e-14, atol=0)
    # including the point itself: first column is the identity
    h0 = g.knn(min(k, 5), exclude_self=False)
    assert np.array_equal(h0[:, 0], np.arange(n))
    assert np.array_equal(wlsqm.knn_hoods(x, k), hoods)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_query_matches_ckdtree(dim):
    n, nq = 15000, 7001
    x = _cloud(n, dim, "uniform", 1)
    rng = np.random.default_rng(2)
    xq = rng.random((nq, dim)) * 1.2 - 0.1           # some queries outside the bounding box
    if dim == 1:
        xq = xq[:, 0].copy()
    tree = cKDTree(x.reshape(n, -1))
    g = wlsqm.PointGrid(x)
    d1, i1 = g.query(xq)
    dr, ir = tree.query(xq.reshape(nq, -1), 1)
    assert i1.dtype == np.int_ and np.array_equal(i1, ir) and np.allclose(d1, dr, rtol=1e-14)
    d5, i5 = g.query(xq, k=5)
    dr5, ir5 = tree.query(xq.reshape(nq, -1), 5)
    assert np.array_equal(i5, ir5) and np.allclose(d5, dr5, rtol=1e-14)
    # a NaN query finds nothing: index n, infinite distance (cKDTree's convention)
    if dim >= 2:
        xb = xq[:4].copy()
        xb[1, 0] = np.nan
        db, ib = g.query(xb)
        assert ib[1] == n and np.isinf(db[1]) and ib[0] == ir[0]


def test_fewer_points_than_k():
    x = _cloud(7, 2, "uniform")
    g = wlsqm.PointGrid(x)
    h = g.knn(10)
    assert (h[:, :6] < 7).all() and (h[:, 6:] == 7).all()


def test_device_tensors_and_gather():
    torch = pytest.importorskip("torch")
    n, k = 30000, 30
    x = _cloud(n, 2, "uniform", 3)
    f = np.sin(x[:, 0]) * np.cos(x[:, 1])
    xt, ft = torch.from_numpy(x).cuda(), torch.from_numpy(f).cuda()
    g = wlsqm.PointGrid(xt)
    hoods = g.knn(k)
    assert hoods.is_cuda and hoods.dtype == torch.int32
    ref = cKDTree(x).query(x, k + 1)[1][:, 1:]
    assert np.array_equal(hoods.cpu().numpy(), ref)
    xk = wlsqm.gather(xt, hoods)
    fk = wlsqm.gather(ft, hoods)
    assert torch.equal(xk, xt[hoods.long()]) and torch.equal(fk, ft[hoods.long()])


@pytest.mark.parametrize("dim,order,k,algo", [(2, 4, 30, 1), (3, 2, 20, 2), (1, 3, 8, 1), (2, 3, 15, 1), (1, 4, 9, 2), (3, 2, 21, 2)])
def test_prepare_solve_hoods_equal_gathered_arrays(dim, order, k, algo):
    """prepare_hoods / solve_hoods == prepare / solve on x[hoods], f[hoods], bit for bit (same kernels, same data)"""
    n = 3001
    x, hoods, f = parity.make_case(n, dim, k)
    no = wlsqm.number_of_dofs(dim, order)
    nk, od, kn, wm = (np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, 1, np.int64),
                      np.full(n, 2, np.int32))
    xk, fk = parity.gathered(x, f, hoods)
    fi0 = np.zeros((n, no))
    fi0[:, 0] = f
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=algo, do_sens=True, max_iter=3)
    s.prepare(x, xk)
    fi_a, sens_a = fi0.copy(), np.zeros((n, k, no))
    it_a = s.solve(fk, fi_a, sens_a)
    s2 = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=algo, do_sens=True, max_iter=3)
    s2.prepare_hoods(x, wlsqm.knn_hoods(x, k))
    fi_b, sens_b = fi0.copy(), np.zeros((n, k, no))
    it_b = s2.solve_hoods(f, fi_b, sens_b)
    assert it_a == it_b
    assert np.array_equal(fi_a, fi_b)
    assert np.array_equal(np.nan_to_num(sens_a), np.nan_to_num(sens_b))


def test_nearest_model_search_on_device_matches_scipy():
    torch = pytest.importorskip("torch")
    n, k, dim, order = 4000, 30, 2, 4
    x, hoods, f = parity.make_case(n, dim, k)
    nk, od, kn, wm = (np.full(n, k, np.int32), np.full(n, order, np.int32), np.zeros(n, np.int64), np.full(n, 1, np.int32))
    xk, fk = parity.gathered(x, f, hoods)
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    s.prepare(x, xk)
    s.solve(fk, np.zeros((n, 15)))
    rng = np.random.default_rng(8)
    xq = x[rng.integers(0, n, 9001)] + 2e-3 * rng.uniform(-1, 1, (9001, 2))
    s.prep_interpolate()                              # SciPy index (reference behaviour)
    out_ref, I_ref = s.interpolate(xq, diff=wlsqm.i2_X)
    s.prep_interpolate(search='gpu')                  # device index
    out_gpu, I_gpu = s.interpolate(xq, diff=wlsqm.i2_X)
    assert I_gpu.dtype == np.int_ and np.array_equal(I_gpu, I_ref) and np.array_equal(out_gpu, out_ref)
    # queries resident on the device: the index stays there too
    out_t, I_t = s.interpolate(torch.from_numpy(xq).cuda(), diff=wlsqm.i2_X)
    torch.cuda.synchronize()
    assert np.array_equal(I_t.cpu().numpy(), I_ref) and np.array_equal(out_t.cpu().numpy(), out_ref)
