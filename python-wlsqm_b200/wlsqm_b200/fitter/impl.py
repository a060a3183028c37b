"""wlsqm_b200.fitter.impl -- where the reference keeps the fitting arithmetic (``wlsqm/fitter/impl.pyx``).

Cython-level only in the reference (``cdef ... nogil`` functions on ``Case*``); nothing is callable from Python.
The same stages run here as CUDA kernels behind the C ABI (``include/wlsqm_b200.h``):

    make_c_1D/2D/3D (impl.pyx:449-544 / 286-432 / 70-269)   ->  prepare_reg_kernel P1        (csrc/wlsqm_prepare.cu)
    make_A (impl.pyx:566-602)                               ->  prepare_reg_kernel P2 (DMMA)
    preprocess_A (impl.pyx:620-689)                         ->  prepare_reg_kernel P3-P4 (Ruiz scaling, pivoted LU)
    solve / solve_contig (impl.pyx:731-846 / 861-974)       ->  prepare_reg_kernel P5 + solve_kernel (csrc/wlsqm_solve.cu)
    solve_iterative (impl.pyx:986-1083)                     ->  solve_kernel<DIM, ITER = true>

The module exists so that ``from wlsqm.fitter import impl`` keeps working.
"""
__all__ = []
