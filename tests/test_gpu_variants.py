"""GPU parity tests of the kernel variants behind the same API (all through the C ABI):

* the packed solve kernel (several small cases per warp pass): tail packs, mixed knowns patterns with the same
  sizes ("geometry-uniform" batches), strided fk (no bulk copy), sensitivities;
* the register/DMMA prepare kernel: several fits per warp through the row phases with odd case counts,
  all-known cases mixed in, and the shared-memory variant as a second implementation of the same arithmetic;
* the interpolation kernel: staged (transposed) and direct all-slots output, pitched output rows, per-model
  orders, 3D.

Tolerances as in tests/test_gpu_parity.py (tests/parity.py states the criterion).
"""
import os

import numpy as np
import pytest

import parity
import oracle as orc

pytestmark = pytest.mark.gpu

wlsqm = pytest.importorskip("wlsqm_b200")


def _scaled_err(got, ref):
    sc = np.abs(ref).max(axis=0)
    sc[sc == 0] = 1.0
    return np.abs(got - ref) / sc


def _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0, do_sens=False, algorithm=1, max_iter=3):
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=algorithm, do_sens=do_sens, max_iter=max_iter)
    s.prepare(x, xk)
    fi = fi0.copy()
    sens = np.zeros((len(nk), xk.shape[1], fi.shape[1])) if do_sens else None
    s.solve(fk, fi, sens)
    return fi, sens, s


PACK_CASES = [
    # dim, order, k, knowns patterns (same popcount), n
    pytest.param(1, 3, 8, (0b0001, 0b0010), 4001, id="1D-o3-k8-F-or-X-known"),
    pytest.param(1, 2, 5, (0,), 1003, id="1D-o2-k5"),
    pytest.param(2, 2, 12, (0b000001, 0b000100, 0b100000), 3001, id="2D-o2-k12-one-known"),
    pytest.param(2, 3, 24, (0b1, 0b100), 2001, id="cfg4-2D-o3-k24-F-or-Y-known"),
    pytest.param(2, 1, 7, (0b011, 0b101), 999, id="2D-o1-k7-two-knowns"),
    pytest.param(3, 1, 9, (0,), 2005, id="3D-o1-k9"),
    pytest.param(3, 2, 21, (0b1, 0b1000000000), 1501, id="3D-o2-k21-one-known"),
]


@pytest.mark.parametrize("dim,order,k,patterns,n", PACK_CASES)
def test_packed_solve_vs_oracle(dim, order, k, patterns, n):
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    no = wlsqm.number_of_dofs(dim, order)
    rng = np.random.default_rng(1)
    kn = np.array(patterns, np.int64)[rng.integers(0, len(patterns), n)]
    nk, od, wm = np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, wlsqm.WEIGHT_CENTER, np.int32)
    fi0 = 0.1 * rng.standard_normal((n, no))
    fi0[:, 0] = f
    fi_g, sens_g, _ = _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0, do_sens=True)
    fi_o, sens_o, _, _ = parity.oracle_solve(dim, nk, od, kn, wm, x, xk, fk, fi0, 1, True)
    for o in range(no):     # known slots are left untouched, bit for bit
        m = (kn >> o & 1).astype(bool)
        assert np.array_equal(fi_g[m, o], fi0[m, o])
    e = _scaled_err(fi_g, fi_o)
    assert np.median(e) < 1e-11 and np.quantile(e, 0.99) < 1e-7, (np.median(e), e.max())
    parity.check_sens(sens_g, sens_o, "packed")


def test_packed_solve_strided_fk_and_device_tensors():
    """fk rows that are not dense (no bulk copy possible) and an fi with extra columns, on the device"""
    torch = pytest.importorskip("torch")
    n, k, dim, order = 2503, 8, 1, 3
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    nk, od, kn, wm = (np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, 1, np.int64),
                      np.full(n, 1, np.int32))
    fi0 = np.zeros((n, 4))
    fi0[:, 0] = f
    fi_ref, _, _ = _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0)
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    s.prepare(torch.from_numpy(x).cuda(), torch.from_numpy(xk).cuda())
    fk_wide = torch.zeros((n, k + 3), dtype=torch.float64, device="cuda")
    fk_wide[:, :k] = torch.from_numpy(fk).cuda()
    fi_wide = torch.full((n, 7), 7.0, dtype=torch.float64, device="cuda")
    fi_wide[:, :4] = torch.from_numpy(fi0).cuda()
    s.solve(fk_wide[:, :k], fi_wide[:, :4])
    torch.cuda.synchronize()
    got = fi_wide.cpu().numpy()
    assert np.array_equal(got[:, :4], fi_ref)
    assert (got[:, 4:] == 7.0).all()


def test_packed_and_one_case_per_warp_kernels_agree():
    """the two solve kernels apply the same operator; only the summation order over the neighbours differs"""
    n, k, dim, order = 3001, 24, 2, 3
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    nk, od, kn, wm = (np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, 1, np.int64),
                      np.full(n, 2, np.int32))
    fi0 = np.zeros((n, 10))
    fi0[:, 0] = f
    fi_pack, _, _ = _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0)
    os.environ["WLSQM_SOLVE_NOPACK"] = "1"
    try:
        fi_warp, _, _ = _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0)
    finally:
        del os.environ["WLSQM_SOLVE_NOPACK"]
    e = _scaled_err(fi_pack, fi_warp)
    # the operator entries of the higher derivatives cancel heavily: re-association noise is cond * eps
    assert np.median(e) < 1e-13 and e.max() < 1e-8, (np.median(e), e.max())


def test_geometry_uniform_batch_large_model():
    """2D order 4 with the known slot varying from case to case: uniform sizes, per-case knowns pattern"""
    n, k, dim, order = 3001, 30, 2, 4
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    rng = np.random.default_rng(2)
    kn = np.array([wlsqm.b2_F, wlsqm.b2_X, wlsqm.b2_Y2], np.int64)[rng.integers(0, 3, n)]
    nk, od, wm = np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, 1, np.int32)
    fi0 = 0.05 * rng.standard_normal((n, 15))
    fi0[:, 0] = f
    for algo in (wlsqm.ALGO_BASIC, wlsqm.ALGO_ITERATIVE):
        fi_g, sens_g, sg = _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0, do_sens=True, algorithm=algo)
        fi_o, sens_o, _, so = parity.oracle_solve(dim, nk, od, kn, wm, x, xk, fk, fi0, algo, True, 3)
        a, b = parity.permuted_self_noise(dim, nk, od, kn, wm, x, xk, fk, fi0, algo, 3)
        for o in range(15):
            m = (kn >> o & 1).astype(bool)
            assert np.array_equal(fi_g[m, o], fi0[m, o])
        # noise-floor criterion of tests/parity.py; known slots hold the caller's values on both sides (error 0)
        print(parity.check_against_floor(fi_g, fi_o, b + (fi_o - a), dim, order, "geometry-uniform algo %d" % algo))
        if algo == wlsqm.ALGO_ITERATIVE:
            # (per-case counts may differ by a round where the bit-exact `norm == prev_norm` exit fires on one side only)
            assert sg.iterations().max() == so.iters.max()
        parity.check_sens(sens_g, sens_o, "geometry-uniform")


@pytest.mark.parametrize("dim,order,k,n", [(2, 4, 30, 1001), (2, 2, 12, 1003), (1, 3, 8, 1002), (3, 1, 10, 1001),
                                            (3, 3, 40, 301), (3, 4, 60, 151)])
def test_prepare_variants_agree_and_odd_counts(dim, order, k, n):
    """register/DMMA kernel (several fits per warp) vs the shared-memory kernel on the same inputs, with a case
    count that leaves the last warp pass partly empty and with all-known cases (silent no-ops) mixed in"""
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    no = wlsqm.number_of_dofs(dim, order)
    nk, od, wm = np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, 2, np.int32)
    kn = np.zeros(n, np.int64)
    kn[3::11] = 1
    kn[::7] = (1 << no) - 1          # everything known: the fit is a no-op
    fi0 = np.zeros((n, no))
    fi0[:, 0] = f
    fi0[::7] = 3.25
    fi_reg, _, _ = _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0)
    os.environ["WLSQM_PREP_KERNEL"] = "smem"
    try:
        fi_smem, _, _ = _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, fi0)
    finally:
        del os.environ["WLSQM_PREP_KERNEL"]
    assert np.array_equal(fi_reg[::7], fi0[::7])
    fi_o, _, _, _ = parity.oracle_solve(dim, nk, od, kn, wm, x, xk, fk, fi0)
    a, b = parity.permuted_self_noise(dim, nk, od, kn, wm, x, xk, fk, fi0)
    parity.check_against_floor(fi_reg, fi_o, b + (fi_o - a), dim, order, "reg-vs-oracle")
    parity.check_against_floor(fi_smem, fi_o, b + (fi_o - a), dim, order, "smem-vs-oracle")


def test_interpolate_variants():
    """all-slots output: staged (dense rows) == direct (pitched rows) ~= one call per slot (few ulp); odd query count"""
    torch = pytest.importorskip("torch")
    for dim, order, k, n in ((2, 4, 30, 700), (3, 3, 40, 300), (1, 4, 9, 500)):
        x, hoods, f = parity.make_case(n, dim, k)
        xk, fk = parity.gathered(x, f, hoods)
        no = wlsqm.number_of_dofs(dim, order)
        nk, od, kn, wm = (np.full(n, k, np.int32), np.full(n, order, np.int32), np.zeros(n, np.int64),
                          np.full(n, 2, np.int32))
        fi_g, _, s = _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, np.zeros((n, no)))
        rng = np.random.default_rng(9)
        nq = 4 * n + 37
        I = rng.integers(0, n, nq).astype(np.int64)
        xq = x[I] + 1e-3 * rng.uniform(-1, 1, (nq,) + x.shape[1:])
        s.prep_interpolate()
        out_all, _ = s.interpolate(xq, diff='all', I=I)
        so = orc.OracleSolver(dim, nk, od, kn, wm)
        so.xi = x
        so.fi = fi_g.copy()
        for d in range(no):
            o1, _ = s.interpolate(xq, diff=d, I=I)
            # all-slots pass = in-place Taylor shift, per-slot call = nested Horner form: same polynomial, two
            # evaluation orders
            assert np.abs(o1 - out_all[:, d]).max() <= 1e-13 * max(np.abs(o1).max(), 1e-300), (dim, d)
            oo = so.interpolate(xq, I, d)
            assert np.abs(o1 - oo).max() <= 1e-12 * max(np.abs(oo).max(), 1e-300), (dim, d)
        # device tensors: dense rows take the staged path, the library's own buffer for host output as well
        xq_t, I_t = torch.from_numpy(xq).cuda(), torch.from_numpy(I).cuda()
        out_t, _ = s.interpolate(xq_t, diff='all', I=I_t)
        torch.cuda.synchronize()
        assert np.array_equal(out_t.cpu().numpy(), out_all)


def test_interpolate_per_model_orders():
    n, k, dim = 1200, 30, 2
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    rng = np.random.default_rng(4)
    od = rng.integers(1, 5, n).astype(np.int32)
    nk, kn, wm = np.full(n, k, np.int32), np.zeros(n, np.int64), np.full(n, 1, np.int32)
    fi_g, _, s = _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, np.zeros((n, 15)))
    nq = 3001
    I = rng.integers(0, n, nq).astype(np.int64)
    xq = x[I] + 1e-3 * rng.uniform(-1, 1, (nq, 2))
    s.prep_interpolate()
    so = orc.OracleSolver(dim, nk, od, kn, wm)
    so.xi = x
    so.fi = fi_g.copy()
    for d in (0, wlsqm.i2_X, wlsqm.i2_XY, wlsqm.i2_Y3, wlsqm.i2_X2Y2):
        og, _ = s.interpolate(xq, diff=d, I=I)
        oo = so.interpolate(xq, I, d)
        assert np.abs(og - oo).max() <= 1e-12 * max(np.abs(oo).max(), 1e-300), d


def test_guest_mode_borrows_the_hosts_operators():
    """ExpertSolver(host=...) (expert.pyx:163-189): same geometry, another field -- the guest owns no operators"""
    n, k, dim, order = 2001, 30, 2, 4
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    g2 = np.cos(3.0 * x[:, 0]) * x[:, 1]
    gk = np.ascontiguousarray(g2[hoods])
    nk, od, kn, wm = (np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, 1, np.int64),
                      np.full(n, 2, np.int32))
    for algo in (wlsqm.ALGO_BASIC, wlsqm.ALGO_ITERATIVE):
        host = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=algo, max_iter=3)
        host.prepare(x, xk)
        guest = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=algo, max_iter=3, do_sens=True, host=host)
        assert guest._borrows
        guest.prepare(x, xk)            # arguments are ignored in guest mode, like the reference
        assert guest.ready and guest.memory_used()[0] < host.memory_used()[0] / 10
        fi_h = np.zeros((n, 15)); fi_h[:, 0] = f
        fi_g = np.zeros((n, 15)); fi_g[:, 0] = g2
        sens = np.zeros((n, k, 15))
        host.solve(fk, fi_h)
        guest.solve(gk, fi_g, sens)
        # an independent solver on the same geometry gives the same numbers, bit for bit
        fi_i, sens_i, _ = _solve_gpu(dim, nk, od, kn, wm, x, xk, gk, np.where(np.arange(15) == 0, g2[:, None], 0.0),
                                     do_sens=True, algorithm=algo)
        assert np.array_equal(fi_g, fi_i)
        assert np.array_equal(np.nan_to_num(sens), np.nan_to_num(sens_i))
        # the host's own solution is not disturbed by the guest
        fi_h2 = np.zeros((n, 15)); fi_h2[:, 0] = f
        host.solve(fk, fi_h2)
        assert np.array_equal(fi_h, fi_h2)
        guest.close()
        host.close()
    with pytest.raises(ValueError, match="must match"):
        host = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
        host.prepare(x, xk)
        wlsqm.ExpertSolver(dim, nk, od, np.zeros(n, np.int64), wm, host=host)


def test_sharded_solver_rows_equal_unsharded_bit_for_bit():
    """the multi-GPU sharding of wlsqm_b200.parallel, exercised on one device: the ranks' contiguous case ranges,
    solved one after the other (gathered arrays and the hoods path), concatenate to the unsharded result exactly"""
    torch = pytest.importorskip("torch")
    from wlsqm_b200 import parallel
    n, k, dim, order = 3001, 30, 2, 4
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    nk, od, kn, wm = (np.full(n, k, np.int32), np.full(n, order, np.int32), np.zeros(n, np.int64),
                      np.full(n, 1, np.int32))
    fi_full, _, _ = _solve_gpu(dim, nk, od, kn, wm, x, xk, fk, np.zeros((n, 15)))
    x_d, f_d = torch.from_numpy(x).cuda(), torch.from_numpy(f).cuda()
    hoods_d = torch.from_numpy(hoods).cuda()
    world = 3
    parts, parts_h = [], []
    for rank in range(world):
        sh = parallel.ShardedExpertSolver(dim, nk, od, kn, wm, rank=rank, world=world)
        sh.prepare(x, xk)
        fi = np.zeros((n, 15))
        sh.solve(fk, fi)
        assert not fi[:sh.lo].any() and not fi[sh.hi:].any()          # other ranks' rows are not touched
        parts.append(fi[sh.lo:sh.hi])
        sh.prepare_hoods(x_d, hoods_d)
        fi_h = torch.zeros((sh.hi - sh.lo, 15), dtype=torch.float64, device="cuda")
        sh.solve_hoods(f_d, fi_h, local=True)
        parts_h.append(fi_h.cpu().numpy())
    assert np.array_equal(np.concatenate(parts), fi_full)
    assert np.array_equal(np.concatenate(parts_h), fi_full)


_SWITCH_SCRIPT = r"""
import sys, numpy as np
sys.path[:0] = [%(root)r, %(root)r + "/python-wlsqm_b200", %(root)r + "/tests", %(root)r + "/oracle"]
import wlsqm_b200 as wlsqm, parity
from wlsqm_b200.utils import lapackdrivers as ld
n, k = 1500, 24
x, hoods, f = parity.make_case(n, 2, k)
xk, fk = parity.gathered(x, f, hoods)
m = (np.full(n, k, np.int32), np.full(n, 3, np.int32), np.full(n, 1, np.int64), np.full(n, 2, np.int32))
fi = np.zeros((n, 10)); fi[:, 0] = f
wlsqm.fit_2D_many_parallel(xk, fk, m[0], x, fi, None, 0, m[1], m[2], m[3])
s = wlsqm.ExpertSolver(2, *m)
s.prepare(x, xk)
fe = np.zeros((n, 10)); fe[:, 0] = f
s.solve(fk, fe)
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((15, 15, 400))); b = np.asfortranarray(rng.standard_normal((15, 400)))
ip = np.zeros((15, 400), np.int32, order="F")
LU = A.copy(order="F"); ld.mgeneralfactor(LU, ip)
xs = b.copy(order="F"); ld.mgeneralfactored(LU, ip, xs)
np.savez(sys.argv[1], fit=fi, expert=fe, LU=LU, ipiv=ip, x=xs)
"""


def test_ab_switches_give_the_same_answers(tmp_path):
    """the A/B switches of the library (memory pool off, one-shot fits through the general path, shared-memory LU,
    no threaded staging) are run in subprocesses and compared with the default configuration"""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    script = tmp_path / "switch.py"
    script.write_text(_SWITCH_SCRIPT % {"root": root})

    def run(env_extra, name):
        env = dict(os.environ)
        env.update(env_extra)
        out = tmp_path / (name + ".npz")
        subprocess.check_call([sys.executable, str(script), str(out)], env=env)
        return np.load(out)
    base = run({}, "base")
    # fused one-shot kernel vs operator path: two orders of the same arithmetic, cond * eps apart
    err = np.abs(base["fit"] - base["expert"]) / np.abs(base["expert"]).max(axis=0)
    assert err.max() <= 1e-7 and np.median(err) <= 1e-12, (err.max(), np.median(err))
    nopool = run({"WLSQM_POOL": "0", "WLSQM_BOUNCE": "0"}, "nopool")
    for key in ("fit", "expert", "LU", "ipiv", "x"):
        assert np.array_equal(base[key], nopool[key]), key
    general = run({"WLSQM_FIT_DIRECT": "0"}, "general")
    assert np.array_equal(general["fit"], general["expert"])          # general one-shot path == ExpertSolver, bit for bit
    assert np.array_equal(general["expert"], base["expert"])
    smem = run({"WLSQM_LU_SMEM": "1"}, "smem")
    assert np.array_equal(smem["ipiv"], base["ipiv"])
    assert np.allclose(smem["LU"], base["LU"], rtol=1e-10, atol=1e-12)
    assert np.allclose(smem["x"], base["x"], rtol=1e-9, atol=1e-11)


ONESHOT_CASES = [
    # dim, order, k, knowns, wm, n
    pytest.param(3, 4, 60, 1, 2, 601, id="3D-o4-k60-bF-center (two matrix rows per lane)"),
    pytest.param(3, 4, 70, 0, 1, 333, id="3D-o4-k70-nr35-uniform"),
    pytest.param(3, 3, 40, 1, 1, 1001, id="3D-o3-k40"),
    pytest.param(2, 4, 30, 0, 1, 2001, id="2D-o4-k30"),
    pytest.param(1, 4, 9, 1, 2, 2003, id="1D-o4-k9"),
]


@pytest.mark.parametrize("dim,order,k,knowns,wm,n", ONESHOT_CASES)
def test_one_shot_kernel_vs_oracle(dim, order, k, knowns, wm, n):
    """fit_*_many_parallel without sens on a uniform batch: ONE fused kernel (assemble, equilibrate, factor, solve; no
    operator is formed) -- generic_fit_basic_many_parallel, simple.pyx:953-1058 -- for every model size including 3D
    order 4 (35 DOFs: two matrix rows per lane, back-substitution through shared memory).  Against the oracle with the
    noise-floor criterion, and against the prepare + solve path."""
    x, hoods, f = parity.make_case(n, dim, k)
    xk, fk = parity.gathered(x, f, hoods)
    no = wlsqm.number_of_dofs(dim, order)
    nk, od, kn, w = np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, knowns, np.int64), np.full(n, wm, np.int32)
    fi0 = np.zeros((n, no)); fi0[:, 0] = f
    got = fi0.copy()
    fit = getattr(wlsqm, "fit_%dD_many_parallel" % dim)
    assert fit(xk, fk, nk, x, got, None, 0, od, kn, w, ntasks=4) == 0
    ref, _, _, _ = parity.oracle_solve(dim, nk, od, kn, w, x, xk, fk, fi0)
    a, b = parity.permuted_self_noise(dim, nk, od, kn, w, x, xk, fk, fi0)
    print(parity.check_against_floor(got, ref, b + (ref - a), dim, order, "one-shot kernel vs oracle"))
    for o in range(no):
        if knowns >> o & 1:
            assert np.array_equal(got[:, o], fi0[:, o])
    exp, _, _ = _solve_gpu(dim, nk, od, kn, w, x, xk, fk, fi0)
    print(parity.check_against_floor(got, exp, b + (exp - a), dim, order, "one-shot kernel vs prepare + solve"))
    # the same through CUDA tensors, bit for bit
    torch = pytest.importorskip("torch")
    got_d = torch.from_numpy(fi0).cuda()
    fit(torch.from_numpy(xk).cuda(), torch.from_numpy(fk).cuda(), nk, torch.from_numpy(x).cuda(), got_d, None, 0, od, kn, w)
    torch.cuda.synchronize()
    assert np.array_equal(got_d.cpu().numpy(), got)
