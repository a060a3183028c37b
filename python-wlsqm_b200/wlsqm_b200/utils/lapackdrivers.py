"""Batched dense drivers on the GPU: ``mgeneral[p]``, ``mgeneralfactor[p]``, ``mgeneralfactored[p]`` and their
symmetric counterparts ``msymmetric[p]``, ``msymmetricfactor[p]``, ``msymmetricfactored[p]``, ``msymmetrize[p]``.

Host-side mirror of the batched families of ``wlsqm/utils/lapackdrivers.pyx`` (general: ``:1551-1723``, on the
fitter's path; symmetric: ``:1107-1354`` and ``:204-278``).  The single-system convenience
wrappers of that module (``general``, ``symmetric``, ``tridiag``, ...) solve one small system per call and have nothing
to batch -- keep importing the reference for those.  The matrix-scaling family (``do_rescale`` and the six
``rescale_*`` algorithms, ``:285-847``) is served as one batched kernel: ``mdo_rescale[p]`` scales nlhs matrices per
launch, and ``do_rescale`` / ``rescale_*`` are the reference's single-matrix entry points on top of it.
Layout is the reference's: ``A`` (n, n, nlhs) Fortran-contiguous, ``b`` (n, nlhs) Fortran,
``ipiv`` (n, nlhs) int32 Fortran, 1-based pivots; everything is overwritten in place; LAPACK's ``info``
is not reported (the reference drops it too: a singular system silently yields inf/NaN).
The ``*p`` variants take ``ntasks`` for signature compatibility; on the GPU every variant is one launch.
"""
from __future__ import annotations

from enum import IntEnum

import numpy as np

from .. import _lib

__all__ = ["ScalingAlgo", "general", "generalfactor", "generalfactored", "symmetric", "symmetricfactor", "symmetricfactored",
           "tridiag", "do_rescale", "mdo_rescale", "mdo_rescalep", "rescale_columns", "rescale_rows", "rescale_twopass",
           "rescale_ruiz2001", "rescale_scalgm", "rescale_dgeequ", "mgeneral", "mgeneralp", "mgeneralfactor", "mgeneralfactorp", "mgeneralfactored", "mgeneralfactoredp",
           "msymmetric", "msymmetricp", "msymmetricfactor", "msymmetricfactorp", "msymmetricfactored",
           "msymmetricfactoredp", "msymmetrize", "msymmetrizep"]


class ScalingAlgo(IntEnum):
    """The reference's enum of matrix-scaling algorithms (``lapackdrivers.pyx:305-317``; plain ints, so comparisons with
    int literals work).  The fitter's ``prepare`` stage uses ``ALGO_RUIZ2001`` (``impl.pyx:620-689``) -- here phase P3 of
    ``prepare_reg_kernel``; ``do_rescale`` / ``mdo_rescale`` below take any of them."""
    ALGO_COLS_EUCL = 1
    ALGO_ROWS_EUCL = 2
    ALGO_TWOPASS = 3
    ALGO_RUIZ2001 = 4
    ALGO_SCALGM = 5
    ALGO_DGEEQU = 6


def _f3(a, name, dtype, ndim):
    """Fortran-contiguous array argument (``double[::1,:,:]`` etc.) -> (pointer, shape, cuda device or None)"""
    dtype = np.dtype(dtype)
    if _lib._is_torch_tensor(a):
        t = a
        if not t.is_cuda:
            raise ValueError(f"{name}: torch tensors must live on a CUDA device (pass numpy arrays for host data)")
        if t.dim() != ndim:
            raise ValueError(f"{name}: Buffer has wrong number of dimensions (expected {ndim}, got {t.dim()})")
        if str(t.dtype).replace("torch.", "") != dtype.name:
            raise ValueError(f"{name}: Buffer dtype mismatch, expected '{dtype.name}' but got '{t.dtype}'")
        exp, acc = [], 1
        for d in t.shape:
            exp.append(acc)
            acc *= int(d)
        # (the stride of an axis of length 1 is arbitrary)
        if any(int(d) > 1 and e != int(st) for d, e, st in zip(t.shape, exp, t.stride())):
            raise ValueError(f"{name}: tensor must be Fortran-contiguous")
        return int(t.data_ptr()), tuple(int(d) for d in t.shape), t.device.index
    if not isinstance(a, np.ndarray):
        raise TypeError(f"{name}: expected a numpy array or a CUDA torch.Tensor")
    if a.dtype != dtype:
        raise ValueError(f"{name}: Buffer dtype mismatch, expected '{dtype.name}' but got '{a.dtype.name}'")
    if a.ndim != ndim:
        raise ValueError(f"{name}: Buffer has wrong number of dimensions (expected {ndim}, got {a.ndim})")
    if not a.flags.f_contiguous:
        raise ValueError(f"{name}: ndarray is not Fortran contiguous")
    if not a.flags.writeable:
        raise ValueError(f"{name}: buffer source array is read-only")
    return a.ctypes.data, a.shape, None


def _dev(*devs):
    """device of the call (first one named), announcing the caller's CUDA stream on it to the library"""
    dev = next((d for d in devs if d is not None), None)
    if dev is None:
        dev = _lib.default_device()
    _lib.announce_stream(dev)
    return dev


def mgeneralfactor(A, ipiv, device=None):
    """LU-factor nlhs independent n x n systems in place (dgetrf each; ``lapackdrivers.pyx:1612-1635``)."""
    pa, sa, da = _f3(A, "A", np.float64, 3)
    pp, sp, dp = _f3(ipiv, "ipiv", np.int32, 2)
    n, n2, nlhs = sa
    if n != n2 or sp != (n, nlhs):
        raise ValueError("shape mismatch: A (n,n,nlhs), ipiv (n,nlhs)")
    _lib.check(_lib.lib().wlsqm_mgetrf(n, nlhs, pa, pp, int(_dev(device, da, dp))))


def mgeneralfactored(LU, ipiv, b, device=None):
    """Solve with factors from :func:`mgeneralfactor`; b is overwritten by x (dgetrs each;
    ``lapackdrivers.pyx:1638-1665``)."""
    pa, sa, da = _f3(LU, "LU", np.float64, 3)
    pp, sp, dp = _f3(ipiv, "ipiv", np.int32, 2)
    pb, sb, db = _f3(b, "b", np.float64, 2)
    n, n2, nlhs = sa
    if n != n2 or sp != (n, nlhs) or sb != (n, nlhs):
        raise ValueError("shape mismatch: LU (n,n,nlhs), ipiv (n,nlhs), b (n,nlhs)")
    _lib.check(_lib.lib().wlsqm_mgetrs(n, nlhs, pa, pp, pb, int(_dev(device, da, dp, db))))


def mgeneral(A, b, device=None):
    """Solve nlhs independent systems (dgesv each; ``lapackdrivers.pyx:1551-1578``): A is overwritten by
    its LU factors, b by the solution."""
    pa, sa, da = _f3(A, "A", np.float64, 3)
    pb, sb, db = _f3(b, "b", np.float64, 2)
    n, n2, nlhs = sa
    if n != n2 or sb != (n, nlhs):
        raise ValueError("shape mismatch: A (n,n,nlhs), b (n,nlhs)")
    ipiv = np.empty((n, nlhs), dtype=np.int32, order='F')
    if da is not None:
        import torch
        ipiv_t = torch.empty((nlhs, n), dtype=torch.int32, device=A.device).t()
        pp = int(ipiv_t.data_ptr())
    else:
        pp = ipiv.ctypes.data
    _lib.check(_lib.lib().wlsqm_mgesv(n, nlhs, pa, pp, pb, int(_dev(device, da, db))))


def mgeneralp(A, b, ntasks=1, device=None):
    """``lapackdrivers.pyx:1581-1609``"""
    return mgeneral(A, b, device)


def mgeneralfactorp(A, ipiv, ntasks=1, device=None):
    """``lapackdrivers.pyx:1666-1692``"""
    return mgeneralfactor(A, ipiv, device)


def mgeneralfactoredp(LU, ipiv, b, ntasks=1, device=None):
    """``lapackdrivers.pyx:1695-1723``"""
    return mgeneralfactored(LU, ipiv, b, device)


# ---- symmetric family (Bunch-Kaufman U D U^T, LAPACK uplo = 'U') ------------------------------------------------

def msymmetricfactor(A, ipiv, device=None):
    """U D U^T-factor nlhs independent symmetric n x n matrices in place (dsytrf 'U' each;
    ``lapackdrivers.pyx:1199-1233``).  Only the upper triangle of each matrix is read and written; ``ipiv`` is
    dsytrf's pivot array."""
    pa, sa, da = _f3(A, "A", np.float64, 3)
    pp, sp, dp = _f3(ipiv, "ipiv", np.int32, 2)
    n, n2, nlhs = sa
    if n != n2 or sp != (n, nlhs):
        raise ValueError("shape mismatch: A (n,n,nlhs), ipiv (n,nlhs)")
    _lib.check(_lib.lib().wlsqm_msytrf(n, nlhs, pa, pp, int(_dev(device, da, dp))))


def msymmetricfactored(A, ipiv, b, device=None):
    """Solve with factors from :func:`msymmetricfactor`; b is overwritten by x (dsytrs 'U' each;
    ``lapackdrivers.pyx:1236-1272``)."""
    pa, sa, da = _f3(A, "A", np.float64, 3)
    pp, sp, dp = _f3(ipiv, "ipiv", np.int32, 2)
    pb, sb, db = _f3(b, "b", np.float64, 2)
    n, n2, nlhs = sa
    if n != n2 or sp != (n, nlhs) or sb != (n, nlhs):
        raise ValueError("shape mismatch: A (n,n,nlhs), ipiv (n,nlhs), b (n,nlhs)")
    _lib.check(_lib.lib().wlsqm_msytrs(n, nlhs, pa, pp, pb, int(_dev(device, da, dp, db))))


def msymmetric(A, b, device=None):
    """Solve nlhs independent symmetric systems (dsysv 'U' each; ``lapackdrivers.pyx:1107-1150``): the upper
    triangle of A is overwritten by its factors, b by the solution."""
    pa, sa, da = _f3(A, "A", np.float64, 3)
    pb, sb, db = _f3(b, "b", np.float64, 2)
    n, n2, nlhs = sa
    if n != n2 or sb != (n, nlhs):
        raise ValueError("shape mismatch: A (n,n,nlhs), b (n,nlhs)")
    _lib.check(_lib.lib().wlsqm_msysv(n, nlhs, pa, None, pb, int(_dev(device, da, db))))


def msymmetrize(A, device=None):
    """``A[:, :, k] = 0.5 * (A[:, :, k] + A[:, :, k].T)`` for every k, in place (``lapackdrivers.pyx:204-230``)."""
    pa, sa, da = _f3(A, "A", np.float64, 3)
    n, n2, nlhs = sa
    if n != n2:
        raise ValueError("shape mismatch: A (n,n,nlhs)")
    _lib.check(_lib.lib().wlsqm_msymmetrize(n, nlhs, pa, int(_dev(device, da))))


def msymmetricp(A, b, ntasks=1, device=None):
    """``lapackdrivers.pyx:1153-1196``"""
    return msymmetric(A, b, device)


def msymmetricfactorp(A, ipiv, ntasks=1, device=None):
    """``lapackdrivers.pyx:1275-1314``"""
    return msymmetricfactor(A, ipiv, device)


def msymmetricfactoredp(A, ipiv, b, ntasks=1, device=None):
    """``lapackdrivers.pyx:1317-1354``"""
    return msymmetricfactored(A, ipiv, b, device)


def msymmetrizep(A, ntasks=1, device=None):
    """``lapackdrivers.pyx:233-278``"""
    return msymmetrize(A, device)


# ---- matrix scaling (equilibration): lapackdrivers.pyx:285-847 -------------------------------------------------------
def mdo_rescale(A, algo, device=None):
    """``do_rescale`` (``lapackdrivers.pyx:319-385``) for nlhs independent matrices in one launch (batched extension in the
    reference's ``m*`` naming): A (nrows, ncols, nlhs) Fortran-contiguous is scaled in place; returns
    ``(row_scale, col_scale)`` of shapes (nrows, nlhs) / (ncols, nlhs), Fortran order -- numpy arrays, or CUDA tensors if A
    is one.  ``scaled_b = b * row_scale``, ``x = scaled_x * col_scale``.  ``algo``: a :class:`ScalingAlgo` value.
    ``LinAlgError`` if ALGO_DGEEQU meets a zero row or column in any matrix (those matrices are left unscaled)."""
    algo = int(algo)
    if algo not in tuple(int(a) for a in ScalingAlgo):
        raise ValueError("Unknown algorithm identifier, got %d" % (algo))
    pa, sa, da = _f3(A, "A", np.float64, 3)
    nrows, ncols, nlhs = sa
    dev = _dev(device, da)
    if da is not None:
        import torch
        rs = torch.empty((nlhs, nrows), dtype=torch.float64, device=A.device).t()
        cs = torch.empty((nlhs, ncols), dtype=torch.float64, device=A.device).t()
        ok = torch.ones((nlhs,), dtype=torch.int32, device=A.device)
        pr, pc, pk = int(rs.data_ptr()), int(cs.data_ptr()), int(ok.data_ptr())
    else:
        rs = np.empty((nrows, nlhs), dtype=np.float64, order="F")
        cs = np.empty((ncols, nlhs), dtype=np.float64, order="F")
        ok = np.ones((nlhs,), dtype=np.int32)
        pr, pc, pk = rs.ctypes.data, cs.ctypes.data, ok.ctypes.data
    _lib.check(_lib.lib().wlsqm_mrescale(nrows, ncols, nlhs, pa, algo, pr, pc, pk, int(dev)))
    bad = (ok == 0)
    if bool(bad.any()):
        idx = np.nonzero(bad.cpu().numpy() if da is not None else bad)[0]
        raise np.linalg.LinAlgError("Matrix scaling failed (e.g. singular row or column) for %d of %d matrices, first: %d"
                                    % (len(idx), nlhs, int(idx[0])))
    return rs, cs


def mdo_rescalep(A, algo, ntasks=1, device=None):
    """``mdo_rescale`` with the ``*p`` signature (``ntasks`` has no meaning on the GPU)"""
    return mdo_rescale(A, algo, device)


def do_rescale(A, algo, device=None):
    """Generic dispatcher for matrix scaling (preconditioning) routines (``lapackdrivers.pyx:319-385``): scales the
    Fortran-contiguous (nrows, ncols) matrix A in place and returns ``(row_scale, col_scale)``."""
    algo = int(algo)
    if algo not in tuple(int(a) for a in ScalingAlgo):
        raise ValueError("Unknown algorithm identifier, got %d" % (algo))
    pa, sa, da = _f3(A, "A", np.float64, 2)
    if da is not None:
        rs, cs = mdo_rescale(A.unsqueeze(2), algo, device)
    else:
        rs, cs = mdo_rescale(A.reshape(sa[0], sa[1], 1, order="F"), algo, device)     # a view: A itself is scaled
    return rs[:, 0], cs[:, 0]


def rescale_columns(A, device=None):
    """column scaling only, Euclidean norm (``lapackdrivers.pyx:394-424``)"""
    return do_rescale(A, ScalingAlgo.ALGO_COLS_EUCL, device)


def rescale_rows(A, device=None):
    """row scaling only, Euclidean norm (``lapackdrivers.pyx:427-453``)"""
    return do_rescale(A, ScalingAlgo.ALGO_ROWS_EUCL, device)


def rescale_twopass(A, device=None):
    """columns first, then rows (``lapackdrivers.pyx:456-495``)"""
    return do_rescale(A, ScalingAlgo.ALGO_TWOPASS, device)


def rescale_dgeequ(A, device=None):
    """LAPACK's DGEEQU (``lapackdrivers.pyx:498-523``)"""
    return do_rescale(A, ScalingAlgo.ALGO_DGEEQU, device)


def rescale_ruiz2001(A, device=None):
    """simultaneous row and column scaling of Ruiz (2001), symmetry-preserving (``lapackdrivers.pyx:526-623``)"""
    return do_rescale(A, ScalingAlgo.ALGO_RUIZ2001, device)


def rescale_scalgm(A, device=None):
    """SCALGM of Chiang and Chandler (2008) (``lapackdrivers.pyx:626-847``)"""
    return do_rescale(A, ScalingAlgo.ALGO_SCALGM, device)


# ---- the reference's single-system entry points, on top of the batched kernels (a batch of one) ------------------------
def _as_batch(A, b=None):
    """(n, n) Fortran matrix [and (n,) vector] as views of shape (n, n, 1) [(n, 1)]: the batched kernels work in place"""
    _f3(A, "A", np.float64, 2)
    A3 = A.unsqueeze(2) if _lib._is_torch_tensor(A) else A.reshape(A.shape[0], A.shape[1], 1, order="F")
    if b is None:
        return A3
    if _lib._is_torch_tensor(b):
        if b.dim() != 1:
            raise ValueError("b: Buffer has wrong number of dimensions (expected 1, got %d)" % b.dim())
        return A3, b.unsqueeze(1)
    b = np.asarray(b) if not isinstance(b, np.ndarray) else b
    if b.ndim != 1 or b.dtype != np.float64 or not b.flags.c_contiguous:
        raise ValueError("b: a contiguous float64 vector is required")
    return A3, b.reshape(b.shape[0], 1, order="F")


def general(A, b, device=None):
    """Solve a general system in place (dgesv; ``lapackdrivers.pyx:1395-1412``): A is destroyed, b becomes the solution."""
    A3, b2 = _as_batch(A, b)
    mgeneral(A3, b2, device)


def generalfactor(A, device=None):
    """LU-factor in place, return the pivots (dgetrf; ``lapackdrivers.pyx:1415-1434``)."""
    A3 = _as_batch(A)
    n = A3.shape[0]
    if _lib._is_torch_tensor(A):
        import torch
        ipiv = torch.empty((1, n), dtype=torch.int32, device=A.device).t()
    else:
        ipiv = np.empty((n, 1), dtype=np.int32, order="F")
    mgeneralfactor(A3, ipiv, device)
    return ipiv[:, 0]


def generalfactored(LU, ipiv, b, device=None):
    """Solve with factors from ``generalfactor`` (dgetrs; ``lapackdrivers.pyx:1437-1462``)."""
    A3, b2 = _as_batch(LU, b)
    mgeneralfactored(A3, ipiv.reshape(-1, 1) if not _lib._is_torch_tensor(ipiv) else ipiv.unsqueeze(1), b2, device)


def symmetric(A, b, device=None):
    """Solve a symmetric system in place, upper triangle (dsysv; ``lapackdrivers.pyx:918-953``)."""
    A3, b2 = _as_batch(A, b)
    msymmetric(A3, b2, device)


def symmetricfactor(A, device=None):
    """U D U^T-factor in place, return the pivots (dsytrf; ``lapackdrivers.pyx:956-1010``)."""
    A3 = _as_batch(A)
    n = A3.shape[0]
    if _lib._is_torch_tensor(A):
        import torch
        ipiv = torch.empty((1, n), dtype=torch.int32, device=A.device).t()
    else:
        ipiv = np.empty((n, 1), dtype=np.int32, order="F")
    msymmetricfactor(A3, ipiv, device)
    return ipiv[:, 0]


def symmetricfactored(A, ipiv, b, device=None):
    """Solve with factors from ``symmetricfactor`` (dsytrs; ``lapackdrivers.pyx:1013-1050``)."""
    A3, b2 = _as_batch(A, b)
    msymmetricfactored(A3, ipiv.reshape(-1, 1) if not _lib._is_torch_tensor(ipiv) else ipiv.unsqueeze(1), b2, device)


def tridiag(a, b, c, x, device=None):
    """The reference's minimal example driver (``lapackdrivers.pyx:854-877``): LAPACK's DGTSV with one right-hand side on
    (DL, D, DU, B) = (a, b, c, x); x becomes the solution, a / b / c are overwritten by the factorisation."""
    arrs = []
    for nm, v in (("a", a), ("b", b), ("c", c), ("x", x)):
        if _lib._is_torch_tensor(v):
            if not v.is_cuda or v.dtype != _torch_f64() or v.dim() != 1 or (v.shape[0] > 1 and v.stride(0) != 1):
                raise ValueError(f"{nm}: a contiguous float64 CUDA vector is required")
            arrs.append((int(v.data_ptr()), v.shape[0], v.device.index))
        else:
            if not isinstance(v, np.ndarray) or v.dtype != np.float64 or v.ndim != 1 or not v.flags.c_contiguous:
                raise ValueError(f"{nm}: Buffer dtype mismatch or wrong layout (a contiguous float64 vector is required)")
            if not v.flags.writeable:
                raise ValueError(f"{nm}: buffer source array is read-only")
            arrs.append((v.ctypes.data, v.shape[0], None))
    n = arrs[1][1]
    if any(m < n for _, m, _ in arrs):
        raise ValueError("a, b, c and x must have at least n = len(b) entries")
    _lib.check(_lib.lib().wlsqm_gtsv(n, arrs[0][0], arrs[1][0], arrs[2][0], arrs[3][0], int(_dev(device, *(d for _, _, d in arrs)))))
    return 0


def _torch_f64():
    import torch
    return torch.float64
