"""The simple API: one-shot local fits, ``fit_{1,2,3}D[_iterative][_many[_parallel]]``.

Host-side mirror of the 18 Python entry points of ``wlsqm/fitter/simple.pyx:60-604`` (same names,
positional order, defaults, in-place ``fi`` / ``sens`` semantics and return values), forwarding to
``wlsqm_fit_many`` of the C ABI (``include/wlsqm_b200.h``), i.e. assemble + factor + solve on the GPU
in one call with the state discarded afterwards (``generic_fit_*``, ``simple.pyx:620-1170``).

``ntasks`` is accepted for signature compatibility and otherwise ignored (the whole batch is one
launch); ``debug`` is accepted and ignored like in the reference's many-case drivers, whose
condition numbers are never returned to the caller.  Every function takes an optional trailing
``device=`` keyword (extension).  Array arguments may be numpy arrays or CUDA ``torch.Tensor``s.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import defs
from .. import _lib

__all__ = [f"fit_{d}D{it}{many}" for d in (1, 2, 3) for it in ("", "_iterative")
           for many in ("", "_many", "_many_parallel")]


def _expand_single(a):
    """view of a single-case array as a batch of one (no copy, so in-place updates reach the caller)"""
    if _lib._is_torch_tensor(a):
        return a.unsqueeze(0)
    return np.asarray(a)[np.newaxis, ...]


def _fit_many(dim, xk, fk, nk, xi, fi, sens, do_sens, order, knowns, weighting_method, algorithm, max_iter, device):
    nk_a = _lib.meta_array(nk, np.int32, "nk")
    order_a = _lib.meta_array(order, np.int32, "order")
    knowns_a = _lib.meta_array(knowns, np.int64, "knowns")
    wm_a = _lib.meta_array(weighting_method, np.int32, "weighting_method")
    ncases = nk_a.shape[0]
    if order_a.shape[0] != ncases or knowns_a.shape[0] != ncases or wm_a.shape[0] != ncases:
        raise ValueError("nk, order, knowns and weighting_method must have the same length")
    if max_iter is None or do_sens is None:      # typed `int` arguments of the reference: Cython raises TypeError
        raise TypeError("do_sens / max_iter: an integer is required")
    if ncases < 1:         # CaseManager_new (infra.pyx:308-360) refuses an empty batch
        raise ValueError("Must specify max_cases > 0 when creating a CaseManager.")
    do_sens = int(do_sens)
    if dim >= 2:
        xk_a = _lib.as_arr(xk, np.float64, 3, "xk")
        xi_a = _lib.as_arr(xi, np.float64, 2, "xi")
        if xi_a.shape[1] < dim or xk_a.shape[2] < dim:
            raise ValueError("xi and xk must have %d coordinates on their last axis" % dim)
        # the library wants dense neighbour rows of exactly `dim` coordinates from host memory; the reference's
        # double[:,:,::contiguous] view also takes pitched rows and longer last axes (only the first dim columns are read)
        if not xk_a.is_cuda and xk_a.shape[1] > 1 and (xk_a.strides[1] != dim or xk_a.shape[2] != dim):
            xk_a = _lib.as_arr(np.ascontiguousarray(xk_a.np[:, :, :dim]), np.float64, 3, "xk")
    else:
        xk_a = _lib.as_arr(xk, np.float64, 2, "xk", last_contig=False, allow_copy=True)
        xi_a = _lib.as_arr(xi, np.float64, 1, "xi", last_contig=False)
        if not xk_a.is_cuda and xk_a.shape[1] > 1 and xk_a.strides[1] != 1:
            xk_a = _lib.as_arr(np.ascontiguousarray(xk_a.np), np.float64, 2, "xk")
    fk_a = _lib.as_arr(fk, np.float64, 2, "fk", last_contig=False, allow_copy=True)
    if not fk_a.is_cuda and fk_a.shape[1] > 1 and fk_a.strides[1] != 1:
        fk_a = _lib.as_arr(np.ascontiguousarray(fk_a.np), np.float64, 2, "fk")
    fi_a = _lib.as_arr(fi, np.float64, 2, "fi", writable=True)
    maxnk, min_order, max_order, _ = _lib.meta_summary(nk_a, order_a, knowns_a, wm_a)
    if ncases and (min_order < 0 or max_order > 4):
        raise ValueError("order must be 0, 1, 2, 3 or 4")
    maxno = defs.NUMBER_OF_DOFS[dim][max_order] if ncases else 1
    for nm, a in (("xk", xk_a), ("fk", fk_a), ("xi", xi_a), ("fi", fi_a)):
        if a.shape[0] < ncases:
            raise ValueError("%s must have at least ncases = %d rows" % (nm, ncases))
    if ncases and (xk_a.shape[1] < maxnk or fk_a.shape[1] < maxnk or fi_a.shape[1] < maxno):
        raise ValueError("xk/fk need >= max(nk) = %d columns and fi >= %d columns" % (maxnk, maxno))
    sens_p, s0, s1 = None, 0, 0
    if do_sens:
        if sens is None:
            raise ValueError("sens must be given when do_sens is set")
        sens_a = _lib.as_arr(sens, np.float64, 3, "sens", writable=True)
        if sens_a.shape[0] < ncases or (ncases and (sens_a.shape[1] < maxnk or sens_a.shape[2] < maxno)):
            raise ValueError("sens must have shape (>= ncases, >= max nk, >= max no)")
        sens_p, s0, s1 = sens_a.ptr, sens_a.strides[0], sens_a.strides[1]
    if device is None:
        device = next((a.device for a in (xk_a, fk_a, fi_a) if a.is_cuda), None)
        if device is None:
            device = _lib.default_device()
    it = C.c_int32(0)
    _lib.announce_stream(device)
    _lib.check(_lib.lib().wlsqm_fit_many(
        dim, ncases, xk_a.ptr, xk_a.strides[0], xk_a.strides[1], fk_a.ptr, fk_a.strides[0], fk_a.strides[1],
        nk_a.ctypes.data, xi_a.ptr, xi_a.strides[0], fi_a.ptr, fi_a.strides[0], sens_p, s0, s1, do_sens,
        order_a.ctypes.data, knowns_a.ctypes.data, wm_a.ctypes.data, algorithm, int(max_iter), int(device),
        C.byref(it)))
    return int(it.value)


def _fit_one(dim, xk, fk, xi, fi, sens, do_sens, order, knowns, weighting_method, algorithm, max_iter, device):
    for nm, v in (("do_sens", do_sens), ("order", order), ("knowns", knowns),
                  ("weighting_method", weighting_method), ("max_iter", max_iter)):
        if v is None:
            raise TypeError("%s: an integer is required" % nm)
    if dim >= 2:
        xk_a = _lib.as_arr(xk, np.float64, 2, "xk")
        xi_b = _expand_single(_lib.as_arr(xi, np.float64, 1, "xi").keep)
    else:
        xk_a = _lib.as_arr(xk, np.float64, 1, "xk", last_contig=False)
        xi_b = np.array([float(xi)], dtype=np.float64)
    _lib.as_arr(fk, np.float64, 1, "fk", last_contig=False)
    _lib.as_arr(fi, np.float64, 1, "fi", writable=True)
    nk = np.array([xk_a.shape[0]], dtype=np.int32)
    sens_b = None
    if int(do_sens):
        if sens is None:
            raise ValueError("sens must be given when do_sens is set")
        _lib.as_arr(sens, np.float64, 2, "sens", writable=True)
        sens_b = _expand_single(sens)
    return _fit_many(dim, _expand_single(xk), _expand_single(fk), nk, xi_b, _expand_single(fi), sens_b, do_sens,
                     np.array([order], dtype=np.int32), np.array([knowns], dtype=np.int64),
                     np.array([weighting_method], dtype=np.int32), algorithm, max_iter, device)


def _make(dim):
    bF = 1  # b{1,2,3}_F

    def fit(xk, fk, xi, fi, sens, do_sens=0, order=2, knowns=bF, weighting_method=defs.WEIGHT_CENTER, debug=0,
            device=None):
        return _fit_one(dim, xk, fk, xi, fi, sens, do_sens, order, knowns, weighting_method, defs.ALGO_BASIC, 0, device)

    def fit_iterative(xk, fk, xi, fi, sens, do_sens=0, order=2, knowns=bF, weighting_method=defs.WEIGHT_CENTER,
                      max_iter=10, debug=0, device=None):
        return _fit_one(dim, xk, fk, xi, fi, sens, do_sens, order, knowns, weighting_method, defs.ALGO_ITERATIVE,
                        max_iter, device)

    def fit_many(xk, fk, nk, xi, fi, sens, do_sens, order, knowns, weighting_method, debug=0, device=None):
        return _fit_many(dim, xk, fk, nk, xi, fi, sens, do_sens, order, knowns, weighting_method, defs.ALGO_BASIC, 0,
                         device)

    def fit_iterative_many(xk, fk, nk, xi, fi, sens, do_sens, order, knowns, weighting_method, max_iter=10, debug=0,
                           device=None):
        return _fit_many(dim, xk, fk, nk, xi, fi, sens, do_sens, order, knowns, weighting_method,
                         defs.ALGO_ITERATIVE, max_iter, device)

    def fit_many_parallel(xk, fk, nk, xi, fi, sens, do_sens, order, knowns, weighting_method, ntasks=8, debug=0,
                          device=None):
        return _fit_many(dim, xk, fk, nk, xi, fi, sens, do_sens, order, knowns, weighting_method, defs.ALGO_BASIC, 0,
                         device)

    def fit_iterative_many_parallel(xk, fk, nk, xi, fi, sens, do_sens, order, knowns, weighting_method, max_iter=10,
                                    ntasks=8, debug=0, device=None):
        return _fit_many(dim, xk, fk, nk, xi, fi, sens, do_sens, order, knowns, weighting_method,
                         defs.ALGO_ITERATIVE, max_iter, device)

    lines = {1: ("simple.pyx:429-478", "481-498", "501-537", "540-559", "562-581", "584-604"),
             2: ("simple.pyx:241-290", "293-310", "318-354", "357-376", "379-398", "401-421"),
             3: ("simple.pyx:60-109", "111-128", "131-167", "170-189", "192-211", "214-234")}[dim]
    out = {}
    for f, suffix, ln, what in (
            (fit, "", lines[0], "Fit one local model"),
            (fit_iterative, "_iterative", lines[1], "Fit one local model, with iterative refinement"),
            (fit_many, "_many", lines[2], "Fit many local models"),
            (fit_iterative_many, "_iterative_many", lines[3], "Fit many local models, with iterative refinement"),
            (fit_many_parallel, "_many_parallel", lines[4], "Fit many local models (one GPU launch; ntasks ignored)"),
            (fit_iterative_many_parallel, "_iterative_many_parallel", lines[5],
             "Fit many local models with iterative refinement (one GPU launch; ntasks ignored)")):
        name = f"fit_{dim}D{suffix}"
        f.__name__ = f.__qualname__ = name
        f.__doc__ = (f"{what} to {dim}D scalar data (reference: wlsqm/fitter/{ln}).\n\n"
                     "fi is updated in place (knowns untouched, unknowns overwritten); sens[k, j] = d fi[j] / d fk[k],\n"
                     "NaN for known j.  Returns the number of refinement iterations taken (0 without _iterative).")
        out[name] = f
    return out


for _d in (1, 2, 3):
    globals().update(_make(_d))
del _d


# ---- the shipped binding is the Cython shim (wlsqm_b200/_shim.pyx); the ctypes functions above are the fallback ----------
import os as _os

BINDING = "ctypes"
if _os.environ.get("WLSQM_BINDING", "cython") != "ctypes":
    try:
        from .. import _shim as _s
        for _n in __all__:
            globals()[_n] = getattr(_s, _n)
        BINDING = _s.BINDING
        del _n, _s
    except ImportError:      # the shim has not been built (python-wlsqm_b200/build_shim.py)
        pass
