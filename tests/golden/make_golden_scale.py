#!/usr/bin/env python
"""Golden vectors for the matrix-scaling family (do_rescale and the six algorithms, lapackdrivers.pyx:285-847) from
the UNMODIFIED reference.   python tests/golden/make_golden_scale.py   -> tests/golden/golden_scale.npz

Batches of small matrices (square Gram-like ones as the fitter produces them, general square, rectangular, badly
scaled); per algorithm the reference's scaled matrix and scale vectors for every matrix of the batch."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "oracle")]
import oracle as orc      # noqa: E402


def batches():
    rng = np.random.default_rng(11)
    out = {}
    for name, (nr, nc, nl) in {"sq3": (3, 3, 40), "gram15": (15, 15, 30), "rect8x5": (8, 5, 25), "rect5x9": (5, 9, 25),
                               "sq36": (36, 36, 6)}.items():
        A = rng.standard_normal((nr, nc, nl))
        if name.startswith("gram"):
            h = 10.0 ** rng.uniform(-3, -1, nl)
            for l in range(nl):                       # normal matrix of monomials up to order 4 at spacing h: badly scaled
                c = np.stack([(h[l] * rng.uniform(-1, 1, 40)) ** (p % 5) * (h[l] * rng.uniform(-1, 1, 40)) ** (p // 5)
                              for p in range(nr)], axis=1)
                A[:, :, l] = c.T @ c
        else:
            A *= 10.0 ** rng.uniform(-4, 4, (nr, 1, nl)) * 10.0 ** rng.uniform(-3, 3, (1, nc, nl))
            A[rng.random((nr, nc, nl)) < 0.15] = 0.0          # sparsity (SCALGM's smallest NON-ZERO magnitude)
            for l in range(nl):                               # but no zero row / column
                for j in range(nr):
                    if not A[j, :, l].any():
                        A[j, rng.integers(nc), l] = 1.0
                for m in range(nc):
                    if not A[:, m, l].any():
                        A[rng.integers(nr), m, l] = 1.0
        out[name] = np.asfortranarray(A)
    return out


def main():
    ref = orc.load_reference()
    if ref is None:
        raise SystemExit("oracle/_ref is not built: run python oracle/build_ref.py first")
    from wlsqm.utils import lapackdrivers as ld
    out = {"names": np.array(sorted(batches()))}
    for name, A in batches().items():
        out[f"{name}/A"] = A
        nr, nc, nl = A.shape
        for algo in range(1, 7):
            S = np.empty_like(A)
            rs, cs = np.empty((nr, nl), order="F"), np.empty((nc, nl), order="F")
            for l in range(nl):
                M = np.asfortranarray(A[:, :, l].copy())
                r, c = ld.do_rescale(M, algo)
                S[:, :, l], rs[:, l], cs[:, l] = M, np.asarray(r), np.asarray(c)
            out[f"{name}/S{algo}"], out[f"{name}/rs{algo}"], out[f"{name}/cs{algo}"] = S, rs, cs
    np.savez_compressed(HERE / "golden_scale.npz", **out)
    print("written", HERE / "golden_scale.npz")


if __name__ == "__main__":
    main()
