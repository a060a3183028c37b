"""Constants of the wlsqm API (ABI: slot order and bit positions are identical to the reference).

Mirrors the Python-visible part of ``wlsqm/fitter/defs.pyx:288-502``: ``ALGO_*``, ``WEIGHT_*``, the DOF
slot indices ``i{1,2,3}_*``, the one-past-end markers ``i?_{1st,2nd,3rd,4th}_end`` (+ ``i3_0th_end``;
the reference does not export ``i1_0th_end`` / ``i2_0th_end``, defs.pyx:316,346,396), ``SIZE{1,2,3}``
and the bitmasks ``b?_* = 1 << i?_*``.  The names are generated from the per-dimension exponent tables
(the same tables the CUDA kernels use, csrc/wlsqm_common.cuh) instead of being written out one by one.
"""

ALGO_BASIC = 1       # defs.pyx:69
ALGO_ITERATIVE = 2   # defs.pyx:70
WEIGHT_UNIFORM = 1   # defs.pyx:74
WEIGHT_CENTER = 2    # defs.pyx:75

#: exponents (a, b, c) of (dx, dy, dz) for every DOF slot, in slot order (defs.pyx:91-183)
SLOT_EXPONENTS = {
    1: [(a, 0, 0) for a in range(5)],
    2: [(d - b, b, 0) for d in range(5) for b in range(d + 1)],
    3: [(0, 0, 0),
        (1, 0, 0), (0, 1, 0), (0, 0, 1),
        (2, 0, 0), (1, 1, 0), (0, 2, 0), (0, 1, 1), (0, 0, 2), (1, 0, 1),
        (3, 0, 0), (2, 1, 0), (1, 2, 0), (0, 3, 0), (0, 2, 1), (0, 1, 2), (0, 0, 3), (1, 0, 2), (2, 0, 1), (1, 1, 1),
        (4, 0, 0), (3, 1, 0), (2, 2, 0), (1, 3, 0), (0, 4, 0), (0, 3, 1), (0, 2, 2), (0, 1, 3), (0, 0, 4), (1, 0, 3),
        (2, 0, 2), (3, 0, 1), (2, 1, 1), (1, 2, 1), (1, 1, 2)],
}

#: number of DOFs per (dimension, order) (infra.pyx:67-112)
NUMBER_OF_DOFS = {1: (1, 2, 3, 4, 5), 2: (1, 3, 6, 10, 15), 3: (1, 4, 10, 20, 35)}


def slot_name(exps) -> str:
    """(2,1,0) -> 'X2Y', (0,0,0) -> 'F'."""
    if not any(exps):
        return "F"
    return "".join(ax + (str(e) if e > 1 else "") for ax, e in zip("XYZ", exps) if e)


def _generate(ns):
    names = []
    for dim, table in SLOT_EXPONENTS.items():
        for idx, exps in enumerate(table):
            nm = slot_name(exps)
            ns[f"i{dim}_{nm}"] = idx
            ns[f"b{dim}_{nm}"] = 1 << idx
            names += [f"i{dim}_{nm}", f"b{dim}_{nm}"]
        ends = NUMBER_OF_DOFS[dim]
        for label, end in zip(("0th", "1st", "2nd", "3rd", "4th"), ends):
            if label == "0th" and dim != 3:
                continue   # not exported by the reference (defs.pyx:316,346)
            ns[f"i{dim}_{label}_end"] = end
            names.append(f"i{dim}_{label}_end")
        ns[f"SIZE{dim}"] = ends[-1]
        names.append(f"SIZE{dim}")
    return names


__all__ = ["ALGO_BASIC", "ALGO_ITERATIVE", "WEIGHT_UNIFORM", "WEIGHT_CENTER"] + _generate(globals())
