"""wlsqm_b200.fitter.polyeval -- where the reference keeps its polynomial evaluators (``wlsqm/fitter/polyeval.pyx``).

Cython-level only in the reference (``taylor_1D/2D/3D``, ``general_1D/2D/3D``: ``cdef ... nogil``); nothing is callable
from Python -- Python callers go through ``interpolate_fit`` / ``ExpertSolver.interpolate``.  Here:

    taylor_* (polyeval.pyx:874-948 / 550-734 / 82-354)    ->  eval_taylor / eval_taylor_nested (csrc/wlsqm_common.cuh)
    general_* (polyeval.pyx:955-1000 / 741-856 / 361-526) ->  the derivative forms of interpolate_kernel (csrc/wlsqm_interp.cu)

The module exists so that ``from wlsqm.fitter import polyeval`` keeps working.
"""
__all__ = []
