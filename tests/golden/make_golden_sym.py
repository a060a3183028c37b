#!/usr/bin/env python
"""Golden vectors for the batched symmetric drivers, produced by the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/build_ref.py): msymmetricfactor / msymmetricfactored / msymmetric / msymmetrize of
wlsqm/utils/lapackdrivers.pyx:204-230,1107-1272.  Run in the build container:  python tests/golden/make_golden_sym.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import oracle as orc  # noqa: E402

ref = orc.load_reference()
assert ref is not None, "build the reference first: python oracle/build_ref.py"
from wlsqm.utils import lapackdrivers as rld  # noqa: E402  (the reference's module, on sys.path after load_reference)

rng = np.random.default_rng(42)
out = {}
names = []
for n, nlhs, kind in ((3, 16, "dense"), (6, 16, "zero_diag"), (15, 12, "dense"), (15, 12, "zero_diag"), (36, 6, "dense")):
    A = rng.standard_normal((n, n, nlhs))
    A = 0.5 * (A + A.transpose(1, 0, 2))
    if kind == "zero_diag":          # forces 2x2 pivot blocks
        for l in range(nlhs):
            A[np.arange(n), np.arange(n), l] = 0.0 if l % 2 == 0 else 1e-3 * rng.standard_normal(n)
    A = np.asfortranarray(A)
    b = np.asfortranarray(rng.standard_normal((n, nlhs)))
    F = A.copy(order="F")
    ipiv = np.zeros((n, nlhs), dtype=np.int32, order="F")
    rld.msymmetricfactor(F, ipiv)
    x = b.copy(order="F")
    rld.msymmetricfactored(F, ipiv, x)
    A2, x2 = A.copy(order="F"), b.copy(order="F")
    rld.msymmetric(A2, x2)
    G = np.asfortranarray(rng.standard_normal((n, n, nlhs)))
    S = G.copy(order="F")
    rld.msymmetrize(S)
    name = f"sym_n{n}_{kind}"
    names.append(name)
    for key, val in (("A", A), ("b", b), ("F", F), ("ipiv", ipiv), ("x", x), ("x_sysv", x2), ("G", G), ("S", S)):
        out[f"{name}/{key}"] = val
out["names"] = np.array(names)
np.savez_compressed(Path(__file__).resolve().parent / "golden_sym.npz", **out)
print("wrote golden_sym.npz:", names, "2x2 blocks present:", any((out[f"{nm}/ipiv"] < 0).any() for nm in names))
