"""ctypes front-end of the C restatement in ``oracle/wlsqm_oracle.c``.

TEST INFRASTRUCTURE ONLY -- the checker for the CUDA path.  Nothing under
``python-wlsqm_b200/`` may import this module (tests/test_no_oracle_in_product.py enforces it).

The classes mirror the reference's Python surface so that parity tests read like the
reference's own tests:

* :class:`OracleSolver`   ~ ``wlsqm.ExpertSolver``      (wlsqm/fitter/expert.pyx:66-781)
* :func:`fit_many`        ~ ``wlsqm.fit_?D_many[_parallel]`` (wlsqm/fitter/simple.pyx:131-604)
* :func:`interpolate_fit` ~ ``wlsqm.interpolate_fit``   (wlsqm/fitter/interp.pyx:34-143)
* :func:`load_reference`  imports the compiled, unmodified reference from ``oracle/_ref`` if present.
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = HERE / "_build" / "libwlsqm_oracle.so"

ALGO_BASIC, ALGO_ITERATIVE = 1, 2
WEIGHT_UNIFORM, WEIGHT_CENTER = 1, 2

_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "wlsqm_oracle.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(HERE), "-s", "-B" if force else "-s"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB))
        _lib.wo_number_of_dofs.restype = C.c_int
        _lib.wo_solve_flat.restype = C.c_int
        _lib.wo_interpolate.restype = C.c_int
        _lib.wo_remap.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def number_of_dofs(dimension: int, order: int) -> int:
    return lib().wo_number_of_dofs(int(dimension), int(order))


def remap(n: int, mask: int):
    o2r = np.empty(n, np.int32)
    r2o = np.empty(n, np.int32)
    nr = lib().wo_remap(_p(o2r), _p(r2o), C.c_int(n), C.c_longlong(int(mask)))
    return nr, o2r, r2o


class OracleSolver:
    """prepare once / solve many, one scalar thread (wlsqm/fitter/expert.pyx:92-655)."""

    def __init__(self, dimension, nk, order, knowns, weighting_method, algorithm=ALGO_BASIC,
                 do_sens=False, max_iter=10):
        self.dim = int(dimension)
        self.nk = np.ascontiguousarray(nk, np.int32)
        self.order = np.ascontiguousarray(order, np.int32)
        self.knowns = np.ascontiguousarray(knowns, np.int64)
        self.wm = np.ascontiguousarray(weighting_method, np.int32)
        self.algorithm, self.do_sens, self.max_iter = int(algorithm), bool(do_sens), int(max_iter)
        n = self.ncases = len(self.nk)
        self.maxnk = int(self.nk.max()) if n else 0
        self.maxno = number_of_dofs(self.dim, int(self.order.max())) if n else 1
        self.c = np.zeros((n, self.maxnk * self.maxno))
        self.w = np.zeros((n, self.maxnk))
        self.LU = np.zeros((n, self.maxno * self.maxno))
        self.As = np.zeros((n, self.maxno * self.maxno))
        self.row = np.zeros((n, self.maxno))
        self.col = np.zeros((n, self.maxno))
        self.ipiv = np.zeros((n, self.maxno), np.int32)
        self.iters = np.zeros(n, np.int32)

    def _pad_xk(self, xk):
        xk = np.asarray(xk, np.float64)
        if self.dim == 1:
            xk = xk.reshape(self.ncases, -1, 1)
        return np.ascontiguousarray(xk[:, :self.maxnk, :])

    def prepare(self, xi, xk):
        self.xi = np.ascontiguousarray(np.asarray(xi, np.float64).reshape(self.ncases, self.dim))
        self.xk = self._pad_xk(xk)
        lib().wo_prepare_flat(self.dim, C.c_long(self.ncases), C.c_long(self.maxnk), self.maxno,
                              _p(self.nk), _p(self.order), _p(self.knowns), _p(self.wm),
                              _p(self.c), _p(self.w), _p(self.LU), _p(self.As), _p(self.row),
                              _p(self.col), _p(self.ipiv), _p(self.xi), _p(self.xk))

    def conds(self):
        """2-norm condition number of the scaled matrix (impl.pyx:662-682)."""
        out = np.full(self.ncases, np.nan)
        for i in range(self.ncases):
            no = number_of_dofs(self.dim, int(self.order[i]))
            nr = no - bin(int(self.knowns[i]) & ((1 << no) - 1)).count("1")
            if nr < 1:
                continue
            s = np.linalg.svd(self.As[i, :nr * nr].reshape(nr, nr), compute_uv=False)
            out[i] = s[0] / s[-1]
        return out

    def solve(self, fk, fi, sens=None):
        """fi is updated in place (first no_j columns of row j), like the reference."""
        fkc = np.ascontiguousarray(np.asarray(fk, np.float64)[:, :self.maxnk])
        fic = np.ascontiguousarray(fi, np.float64)
        sn = None
        if self.do_sens:
            sn = np.ascontiguousarray(sens[:, :self.maxnk, :], np.float64)
            assert sn.shape[2] == fic.shape[1]
        it = lib().wo_solve_flat(self.dim, C.c_long(self.ncases), C.c_long(self.maxnk), self.maxno,
                                 _p(self.nk), _p(self.order), _p(self.knowns), _p(self.wm),
                                 _p(self.c), _p(self.w), _p(self.LU), _p(self.row), _p(self.col),
                                 _p(self.ipiv), _p(fkc), _p(fic), C.c_long(fic.shape[1]), _p(sn),
                                 self.algorithm, self.max_iter, _p(self.xi), _p(self.xk),
                                 _p(self.iters))
        fi[...] = fic
        self.fi = fic.copy()
        if sn is not None:
            sens[:, :self.maxnk, :] = sn
        return it

    def interpolate(self, x, I, diff=0):
        x = np.ascontiguousarray(np.asarray(x, np.float64).reshape(len(I), self.dim))
        I = np.ascontiguousarray(I, np.int64)
        out = np.empty(len(I))
        rc = lib().wo_interpolate(self.dim, _p(self.order), _p(self.xi), _p(self.fi),
                                  C.c_long(self.fi.shape[1]), _p(I), _p(x), C.c_long(len(I)),
                                  int(diff), _p(out))
        if rc:
            raise ValueError("invalid diff")
        return out


    def interpolate_continuous(self, x, r, diff=0):
        """mode='continuous' (expert_interpolate_continuous, wlsqm/fitter/expert.pyx:898-985): for every query, the
        weighted average over all local models whose origin lies within r (cKDTree.query_ball_tree, :902-911), each
        model evaluated by interpolate_nD (:958-972), weights alpha + beta (1 - sqrt(d2 / r^2))^2 with alpha = 0,
        beta = 1 (:45-46, :979-980), accumulated in the order of the ball query's lists (:955-983); cdivision is on
        (:11), so a query with no model within r yields 0/0 = NaN."""
        from scipy.spatial import cKDTree
        x = np.ascontiguousarray(np.asarray(x, np.float64).reshape(-1, self.dim))
        lists = cKDTree(x).query_ball_tree(cKDTree(self.xi), r=r)
        cnt = np.array([len(L) for L in lists], np.int64)
        qm = np.repeat(np.arange(len(x)), cnt)
        li = np.array([i for L in lists for i in L], np.int64)
        out = np.full(len(x), np.nan)
        if len(li) == 0:
            return out
        vals = self.interpolate(x[qm], li, diff)
        d2 = ((x[qm] - self.xi[li]) ** 2).sum(axis=1) if self.dim > 1 else ((x[qm, 0] - self.xi[li, 0]) ** 2)
        tmp = 1.0 - np.sqrt(d2 / (r * r))
        w = 0.0 + 1.0 * tmp * tmp
        acc = np.zeros(len(x))
        sw = np.zeros(len(x))
        # sequential accumulation per query, in list order (np.add.at applies the updates in index order)
        np.add.at(acc, qm, w * vals)
        np.add.at(sw, qm, w)
        with np.errstate(invalid="ignore", divide="ignore"):
            out = acc / sw
        return out


def fit_many(dimension, xk, fk, nk, xi, fi, sens, do_sens, order, knowns, weighting_method,
             algorithm=ALGO_BASIC, max_iter=10):
    """fit_?D[_iterative]_many[_parallel] (wlsqm/fitter/simple.pyx:731-1170): prepare+solve."""
    s = OracleSolver(dimension, nk, order, knowns, weighting_method, algorithm, do_sens, max_iter)
    s.prepare(xi, xk)
    return s.solve(fk, fi, sens)


def interpolate_fit(xi, fi, dimension, order, x, diff=0):
    xi = np.ascontiguousarray(np.atleast_1d(np.asarray(xi, np.float64)))
    fi = np.ascontiguousarray(fi, np.float64)
    x = np.ascontiguousarray(np.asarray(x, np.float64).reshape(-1, dimension))
    out = np.empty(len(x))
    order_a = np.array([order], np.int32)
    rc = lib().wo_interpolate(int(dimension), _p(order_a), _p(xi), _p(fi), C.c_long(len(fi)), None,
                              _p(x), C.c_long(len(x)), int(diff), _p(out))
    if rc:
        raise ValueError("invalid diff")
    return out


def mgetrf(A, ipiv):
    n, _, nlhs = A.shape
    At = np.ascontiguousarray(A.transpose(2, 1, 0))  # [l][m][j] == Fortran (j,m) per system
    pt = np.zeros((nlhs, n), np.int32)
    lib().wo_mgetrf(n, C.c_long(nlhs), _p(At), _p(pt))
    A[...] = At.transpose(2, 1, 0)
    ipiv[...] = pt.T


def mgetrs(LU, ipiv, b):
    n, _, nlhs = LU.shape
    At = np.ascontiguousarray(LU.transpose(2, 1, 0))
    pt = np.ascontiguousarray(ipiv.T, np.int32)
    bt = np.ascontiguousarray(b.T)
    lib().wo_mgetrs(n, C.c_long(nlhs), _p(At), _p(pt), _p(bt))
    b[...] = bt.T


# ---- symmetric indefinite drivers: numpy restatement of LAPACK's dsytf2 / dsytrs, uplo = 'U' ------------------
# The reference reaches them through scipy.linalg.cython_lapack (wlsqm/utils/lapackdrivers.pyx:73-80; call sites
# :1139-1146 dsysv, :1222-1230 dsytrf, :1263-1270 dsytrs), i.e. SciPy's bundled OpenBLAS/LAPACK (SciPy 1.18.1 here;
# the reference pins scipy>=1.9).  The algorithm is the published Bunch-Kaufman diagonal pivoting of LAPACK 3.x
# dsytf2.f / dsytrs.f; pinned by tests/golden/golden_sym.npz (outputs of the unmodified reference).

def sytf2_upper(A):
    """in place on the upper triangle of the (n, n) array A; returns ipiv (int32, LAPACK's 1-based signed convention)"""
    n = A.shape[0]
    ipiv = np.zeros(n, dtype=np.int32)
    alpha = (1.0 + np.sqrt(17.0)) / 8.0
    k = n - 1
    while k >= 0:
        kstep, kp = 1, k
        absakk = abs(A[k, k])
        imax, colmax = 0, 0.0
        if k > 0:
            imax = int(np.argmax(np.abs(A[:k, k])))
            colmax = abs(A[imax, k])
        if max(absakk, colmax) == 0.0 or absakk != absakk:
            kp = k
        else:
            if absakk >= alpha * colmax:
                kp = k
            else:
                rowmax = np.abs(A[imax, imax + 1:k + 1]).max()
                if imax > 0:
                    rowmax = max(rowmax, np.abs(A[:imax, imax]).max())
                if absakk >= alpha * colmax * (colmax / rowmax):
                    kp = k
                elif abs(A[imax, imax]) >= alpha * rowmax:
                    kp = imax
                else:
                    kp, kstep = imax, 2
            kk = k - kstep + 1
            if kp != kk:
                A[:kp, [kk, kp]] = A[:kp, [kp, kk]]
                t = A[kp + 1:kk, kk].copy()
                A[kp + 1:kk, kk] = A[kp, kp + 1:kk]
                A[kp, kp + 1:kk] = t
                A[kk, kk], A[kp, kp] = A[kp, kp], A[kk, kk]
                if kstep == 2:
                    A[k - 1, k], A[kp, k] = A[kp, k], A[k - 1, k]
            if kstep == 1:
                r1 = 1.0 / A[k, k]
                x = A[:k, k].copy()
                for j in range(k):
                    if x[j] != 0.0:
                        A[:j + 1, j] += x[:j + 1] * (-r1 * x[j])
                A[:k, k] *= r1
            elif k > 1:
                d12 = A[k - 1, k]
                d22, d11 = A[k - 1, k - 1] / d12, A[k, k] / d12
                t = 1.0 / (d11 * d22 - 1.0)
                d12 = t / d12
                for j in range(k - 2, -1, -1):
                    wkm1 = d12 * (d11 * A[j, k - 1] - A[j, k])
                    wk = d12 * (d22 * A[j, k] - A[j, k - 1])
                    A[:j + 1, j] = A[:j + 1, j] - A[:j + 1, k] * wk - A[:j + 1, k - 1] * wkm1
                    A[j, k], A[j, k - 1] = wk, wkm1
        if kstep == 1:
            ipiv[k] = kp + 1
        else:
            ipiv[k] = ipiv[k - 1] = -(kp + 1)
        k -= kstep
    return ipiv


def sytrs_upper(A, ipiv, b):
    """b <- A^-1 b with the factors of sytf2_upper, in place"""
    n = A.shape[0]
    k = n - 1
    while k >= 0:
        if ipiv[k] > 0:
            kp = ipiv[k] - 1
            if kp != k:
                b[k], b[kp] = b[kp], b[k]
            b[:k] -= A[:k, k] * b[k]
            b[k] *= 1.0 / A[k, k]
            k -= 1
        else:
            kp = -ipiv[k] - 1
            if kp != k - 1:
                b[k - 1], b[kp] = b[kp], b[k - 1]
            b[:k - 1] -= A[:k - 1, k] * b[k]
            b[:k - 1] -= A[:k - 1, k - 1] * b[k - 1]
            akm1k = A[k - 1, k]
            akm1, ak = A[k - 1, k - 1] / akm1k, A[k, k] / akm1k
            denom = akm1 * ak - 1.0
            bkm1, bk = b[k - 1] / akm1k, b[k] / akm1k
            b[k - 1] = (ak * bkm1 - bk) / denom
            b[k] = (akm1 * bk - bkm1) / denom
            k -= 2
    k = 0
    while k < n:
        if ipiv[k] > 0:
            b[k] -= A[:k, k] @ b[:k]
            kp = ipiv[k] - 1
            if kp != k:
                b[k], b[kp] = b[kp], b[k]
            k += 1
        else:
            b[k] -= A[:k, k] @ b[:k]
            b[k + 1] -= A[:k, k + 1] @ b[:k]
            kp = -ipiv[k] - 1
            if kp != k:
                b[k], b[kp] = b[kp], b[k]
            k += 2
    return b


def msytrf(A, ipiv):
    """batched restatement of msymmetricfactor_c (lapackdrivers.pyx:1215-1233): A (n,n,nlhs) F, ipiv (n,nlhs) F"""
    for l in range(A.shape[2]):
        M = np.array(A[:, :, l])
        ipiv[:, l] = sytf2_upper(M)
        iu = np.triu_indices(A.shape[0])
        A[:, :, l][iu] = M[iu]


def msytrs(A, ipiv, b):
    """batched restatement of msymmetricfactored_c (lapackdrivers.pyx:1263-1272)"""
    for l in range(A.shape[2]):
        x = np.array(b[:, l])
        sytrs_upper(np.array(A[:, :, l]), ipiv[:, l], x)
        b[:, l] = x


def msymmetrize(A):
    """msymmetrize_c (lapackdrivers.pyx:219-230)"""
    n = A.shape[0]
    for j in range(1, n):
        for i in range(j):
            t = 0.5 * (A[i, j, :] + A[j, i, :])
            A[i, j, :] = t
            A[j, i, :] = t


def _cpu_has(*flags):
    try:
        txt = Path("/proc/cpuinfo").read_text()
        line = next(l for l in txt.splitlines() if l.startswith("flags"))
        have = set(line.split(":", 1)[1].split())
        return all(f in have for f in flags)
    except Exception:
        return False


def reference_variants():
    """{name: directory} of the reference builds present AND runnable on this host (oracle/build_ref.py)"""
    out = {}
    if (HERE / "_ref" / "wlsqm" / "__init__.py").exists():
        out["generic"] = HERE / "_ref"
    if (HERE / "_ref" / "_tuned" / "wlsqm" / "__init__.py").exists() and _cpu_has("avx2", "fma", "bmi2"):
        out["tuned"] = HERE / "_ref" / "_tuned"
    return out


def load_reference(variant: str = "generic"):
    """Import the compiled unmodified reference from oracle/_ref (None if it is not built).  One variant per process:
    `variant` = "generic" (-O2, the checker) or "tuned" (-O2 -march=x86-64-v3, CPU-baseline timing only)."""
    ref = reference_variants().get(variant)
    if ref is None:
        return None
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    try:
        mod = importlib.import_module("wlsqm")
    except Exception:  # pragma: no cover - e.g. ABI mismatch on a foreign box
        return None
    if Path(mod.__file__).resolve().parent.parent != Path(ref).resolve():
        return None      # another variant is already imported in this process
    return mod


# ---- matrix scaling: restatement of wlsqm/utils/lapackdrivers.pyx:285-847 (test infrastructure) -----------------------
def do_rescale(A, algo):
    """do_rescale (lapackdrivers.pyx:319-385): scales the (nrows, ncols) matrix A in place, returns (row_scale, col_scale).
    algo: 1 rescale_columns_c (:412-424), 2 rescale_rows_c (:441-453), 3 rescale_twopass_c (:474-495), 4 rescale_ruiz2001_c
    (:553-623), 5 rescale_scalgm_c (:626-847), 6 rescale_dgeequ_c (:517-523 -> LAPACK DGEEQU).  Plain Python loops in the
    reference's order of operations (small matrices only).  Raises numpy.linalg.LinAlgError where the reference does."""
    import math
    nrows, ncols = A.shape
    rs, cs = [1.0] * nrows, [1.0] * ncols
    a = [[float(A[j, m]) for m in range(ncols)] for j in range(nrows)]
    eps = 1e-15                                                                  # lapackdrivers.pyx:87

    def cols_eucl():
        for m in range(ncols):
            c, acc = cs[m], 0.0
            for j in range(nrows):
                tmp = a[j][m] * (c * rs[j])
                acc += tmp * tmp
            cs[m] /= math.sqrt(acc)

    def rows_eucl():
        for j in range(nrows):
            r, acc = rs[j], 0.0
            for m in range(ncols):
                tmp = a[j][m] * (r * cs[m])
                acc += tmp * tmp
            rs[j] /= math.sqrt(acc)

    if algo == 1:
        cols_eucl()
    elif algo == 2:
        rows_eucl()
    elif algo == 3:
        cols_eucl()
        rows_eucl()
    elif algo == 6:
        # DGEEQU (LAPACK 3.x dgeequ.f): smlnum = dlamch('S'), bignum = 1 / smlnum
        sml = sys.float_info.min
        big = 1.0 / sml
        r = [max(abs(a[j][m]) for m in range(ncols)) for j in range(nrows)]
        if min(r) == 0.0:
            raise np.linalg.LinAlgError("Matrix scaling failed (e.g. singular row or column).")
        rs = [1.0 / min(max(v, sml), big) for v in r]
        c = [max(abs(a[j][m]) * rs[j] for j in range(nrows)) for m in range(ncols)]
        if min(c) == 0.0:
            raise np.linalg.LinAlgError("Matrix scaling failed (e.g. singular row or column).")
        cs = [1.0 / min(max(v, sml), big) for v in c]
    elif algo == 4:
        drp, dcp = [1.0] * nrows, [1.0] * ncols
        for _k in range(100):
            dr = [math.sqrt(max([0.0] + [abs(a[j][m] / (drp[j] * dcp[m])) for m in range(ncols)])) for j in range(nrows)]
            dc = [math.sqrt(max([0.0] + [abs(a[j][m] / (dcp[m] * drp[j])) for j in range(nrows)])) for m in range(ncols)]
            for j in range(nrows):
                drp[j] *= dr[j]
                rs[j] /= dr[j]
            for m in range(ncols):
                dcp[m] *= dc[m]
                cs[m] /= dc[m]
            if max(abs(1.0 - v * v) for v in dr) < eps and max(abs(1.0 - v * v) for v in dc) < eps:
                break
    elif algo == 5:
        def up(vals):          # smallest non-zero magnitude (:664-683)
            acc = 0.0
            for tmp in vals:
                if acc == 0.0 or (tmp > 0.0 and tmp < acc):
                    acc = tmp
            return 1.0 / acc if acc != 0.0 else math.inf

        def down(vals):        # largest magnitude (:712-731)
            acc = 0.0
            for tmp in vals:
                if tmp > acc:
                    acc = tmp
            return 1.0 / acc if acc != 0.0 else math.inf

        def rows(f, mod_cs):
            if mod_cs is None:
                return [f([abs(a[j][m] * (rs[j] * cs[m])) for m in range(ncols)]) for j in range(nrows)]
            return [f([abs(a[j][m] * (rs[j] * cs[m] * mod_cs[m])) for m in range(ncols)]) for j in range(nrows)]

        def cols(f, mod_rs):
            if mod_rs is None:
                return [f([abs(a[j][m] * (cs[m] * rs[j])) for j in range(nrows)]) for m in range(ncols)]
            return [f([abs(a[j][m] * (cs[m] * rs[j] * mod_rs[j])) for j in range(nrows)]) for m in range(ncols)]
        mode = 1
        for _k in range(100):
            for f in ((up, down) if mode == 1 else (down,)):
                dr1 = rows(f, None)
                dc1 = cols(f, dr1)
                dc2 = cols(f, None)
                dr2 = rows(f, dc2)
                for j in range(nrows):
                    rs[j] *= math.sqrt(dr1[j] * dr2[j])
                for m in range(ncols):
                    cs[m] *= math.sqrt(dc1[m] * dc2[m])
            er = max(abs(1.0 - max([0.0] + [abs(a[j][m] * (rs[j] * cs[m])) for m in range(ncols)])) for j in range(nrows))
            ec = max(abs(1.0 - max([0.0] + [abs(a[j][m] * (cs[m] * rs[j])) for j in range(nrows)])) for m in range(ncols))
            if er < eps and ec < eps:
                if mode == 1:
                    mode = 2
                else:
                    break
    else:
        raise ValueError("Unknown algorithm identifier, got %d" % (algo))
    for m in range(ncols):                                                        # apply_scaling_c (:293-299)
        for j in range(nrows):
            A[j, m] = a[j][m] * (rs[j] * cs[m])
    return np.array(rs), np.array(cs)
