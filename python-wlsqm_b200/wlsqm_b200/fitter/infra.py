"""wlsqm_b200.fitter.infra -- where the reference keeps its per-case state (``wlsqm/fitter/infra.pyx``).

In the reference this module is Cython-level only (``Case`` / ``CaseManager`` structs and their ``cdef``
functions, ``infra.pxd:124-182``); nothing in it is callable from Python.  Here that state lives on the
device behind the C ABI (``include/wlsqm_b200.h``):

    Case, CaseManager (infra.pyx:308-795)   ->  ``wlsqm_solver`` (csrc/wlsqm_capi.cu): one 32-byte ``CaseMeta`` per case
                                                 (none at all for uniform batches) + one operator block per case
    remap (infra.pyx:145-200)               ->  ``R2O`` table built in ``prepare_reg_kernel`` (csrc/wlsqm_prepare.cu)
    Case_make_weights (infra.pyx:668-702)   ->  phase P1 of ``prepare_reg_kernel``
    Case_set_fi / Case_get_fi (:780-795)    ->  the solver-owned ``fi`` copy written by ``solve_kernel`` (csrc/wlsqm_solve.cu)

The module exists so that ``from wlsqm.fitter import infra`` keeps working; the integer helpers below are the
host-side size arithmetic the Python mirror itself uses (``number_of_dofs``: infra.pyx:67-112).
"""
from .defs import NUMBER_OF_DOFS

__all__ = ["number_of_dofs", "number_of_reduced_dofs"]


def number_of_dofs(dimension, order):
    """DOFs of the full model; -1 for a bad dimension, -2 for a bad order (infra.pyx:67-112)"""
    if dimension not in (1, 2, 3):
        return -1
    if order not in (0, 1, 2, 3, 4):
        return -2
    return NUMBER_OF_DOFS[dimension][order]


def number_of_reduced_dofs(n, knowns):
    """unknown DOFs: n minus the number of set bits among the low n bits of `knowns` (infra.pyx:119-121)"""
    return int(n) - bin(int(knowns) & ((1 << int(n)) - 1)).count("1")
