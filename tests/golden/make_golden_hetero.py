#!/usr/bin/env python
"""Golden vectors for batches with per-case nk / order / knowns / weighting, from the UNMODIFIED reference.

Run in the build container after `python oracle/build_ref.py`:  python tests/golden/make_golden_hetero.py
Inputs are tests/parity.py::hetero_case(dim) (seeded; 3000 cases in 2D and in 3D, orders 0-4, F known or not, a mixed
second derivative known in 30 % of the order >= 2 cases, both weightings) -- the same arrays
tests/test_gpu_parity.py::test_heterogeneous_batch feeds to the GPU.  Stored per dimension: the reference's fi, the
NaN pattern of its sens (packed bits) and sens of every 50th case.  -> tests/golden/golden_hetero.npz (ALGO_BASIC)
and tests/golden/golden_hetero_iter.npz (ALGO_ITERATIVE, max_iter = 3, plus the returned iteration count)
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests"), str(ROOT / "python-wlsqm_b200")]
import oracle as orc      # noqa: E402
import parity             # noqa: E402


def main():
    ref = orc.load_reference()
    if ref is None:
        raise SystemExit("oracle/_ref is not built: run python oracle/build_ref.py first")
    out = {}
    for dim in (2, 3):
        c = parity.hetero_case(dim)
        n, kmax = c["n"], c["kmax"]
        s = ref.ExpertSolver(dim, c["nk"], c["od"], c["kn"], c["wm"], algorithm=ref.ALGO_BASIC, do_sens=True, max_iter=10,
                             ntasks=1)
        s.prepare(c["x"], c["xk"])
        fi = c["fi0"].copy()
        sens = np.zeros((n, kmax, fi.shape[1]))
        s.solve(c["fk"], fi, sens)
        out["d%d/fi_ref" % dim] = fi
        out["d%d/sens_nan_bits" % dim] = np.packbits(np.isnan(sens).ravel())
        out["d%d/sens_ref_every50" % dim] = sens[::50]
        # guards against drift of the seeded inputs
        out["d%d/input_checksum" % dim] = np.array([c["xk"].sum(), c["fk"].sum(), float(c["nk"].sum()), float(c["od"].sum()),
                                                    float(c["kn"].sum()), float(c["wm"].sum()), c["fi0"].sum()])
        print("dim %d: %d cases, fi checksum %.17g" % (dim, n, fi.sum()))
    np.savez_compressed(HERE / "golden_hetero.npz", **out)
    print("written", HERE / "golden_hetero.npz")
    # the same batches through ALGO_ITERATIVE (impl.pyx:986-1083; max_iter = 3, with sens): per-case records AND the
    # refinement loop with its bit-pattern-dependent `norm == prev_norm` exit -> tests/golden/golden_hetero_iter.npz
    out = {}
    for dim in (2, 3):
        c = parity.hetero_case(dim)
        n, kmax = c["n"], c["kmax"]
        s = ref.ExpertSolver(dim, c["nk"], c["od"], c["kn"], c["wm"], algorithm=ref.ALGO_ITERATIVE, do_sens=True, max_iter=3,
                             ntasks=1)
        s.prepare(c["x"], c["xk"])
        fi = c["fi0"].copy()
        sens = np.zeros((n, kmax, fi.shape[1]))
        it = s.solve(c["fk"], fi, sens)
        out["d%d/fi_ref" % dim] = fi
        out["d%d/iters_max" % dim] = np.array([it], np.int32)
        out["d%d/sens_nan_bits" % dim] = np.packbits(np.isnan(sens).ravel())
        out["d%d/sens_ref_every50" % dim] = sens[::50]
        print("dim %d iterative: max iterations %d, fi checksum %.17g" % (dim, it, fi.sum()))
    np.savez_compressed(HERE / "golden_hetero_iter.npz", **out)
    print("written", HERE / "golden_hetero_iter.npz")


if __name__ == "__main__":
    main()
