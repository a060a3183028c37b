#!/usr/bin/env python
"""Golden vectors for ExpertSolver.interpolate (both modes) from the UNMODIFIED reference.

Run in the build container after `python oracle/build_ref.py`:  python tests/golden/make_golden_continuous.py

The reference's own tests never call ExpertSolver.prep_interpolate / interpolate, so these
fixtures are what pins mode='continuous' (expert_interpolate_continuous, wlsqm/fitter/expert.pyx:898-985) and the
ExpertSolver front end of mode='nearest' (:830-895).  Per dimension: a seeded cloud (tests/parity.py::make_case), the
reference's fit `fi`, queries near the cloud plus a few far outside it (no model within r: the reference returns
0/0 = NaN, cdivision on), the averaging radius r, and the reference's outputs for the function value and two
derivative slots in both modes, with the nearest-model index I_out.  -> tests/golden/golden_continuous.npz
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests"), str(ROOT / "python-wlsqm_b200")]
import oracle as orc      # noqa: E402
import parity             # noqa: E402

# dim -> (n, k, order, knowns, weighting, r in units of the point spacing h0, derivative slots)
CASES = {1: (400, 6, 3, 0, 2, 2.5, (0, 1, 3)),
         2: (1500, 20, 3, 0, 2, 3.0, (0, 1, 4)),      # i2_F, i2_X, i2_XY
         3: (1200, 30, 2, 1, 1, 3.5, (0, 3, 5))}      # i3_F, i3_Z, i3_XY
NQ, NFAR = 300, 6


def inputs(dim):
    """seeded inputs, shared with the tests (tests/test_oracle_vs_golden.py, tests/test_gpu_golden.py)"""
    import workloads as wl
    n, k, order, knowns, wm, rh, diffs = CASES[dim]
    x, hoods, f = parity.make_case(n, dim, k, seed=100 + dim)
    xk, fk = parity.gathered(x, f, hoods)
    no = orc.number_of_dofs(dim, order)
    fi0 = np.zeros((n, no))
    fi0[:, 0] = f
    rng = np.random.default_rng(200 + dim)
    x2 = x.reshape(n, -1)
    xq = x2[rng.integers(0, n, NQ)] + 0.4 * wl.H0 * rng.uniform(-1, 1, (NQ, dim))
    xq[-NFAR:] += 1000.0 * wl.H0 * n ** (1.0 / dim)          # far outside the cloud: no model within r
    if dim == 1:
        xq = np.ascontiguousarray(xq[:, 0])
    meta = (np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, knowns, np.int64), np.full(n, wm, np.int32))
    return dict(dim=dim, n=n, k=k, order=order, no=no, x=x, xk=xk, fk=fk, fi0=fi0, xq=xq, r=rh * wl.H0, diffs=diffs, meta=meta)


def main():
    ref = orc.load_reference()
    if ref is None:
        raise SystemExit("oracle/_ref is not built: run python oracle/build_ref.py first")
    out = {}
    for dim in (1, 2, 3):
        c = inputs(dim)
        s = ref.ExpertSolver(dim, *c["meta"], algorithm=ref.ALGO_BASIC, do_sens=False, ntasks=1)
        s.prepare(c["x"], c["xk"])
        fi = c["fi0"].copy()
        s.solve(c["fk"], fi)
        s.prep_interpolate()
        out["d%d/fi_ref" % dim] = fi
        for d in c["diffs"]:
            oc, Ic = s.interpolate(c["xq"], mode="continuous", r=c["r"], diff=d)
            if dim == 1:
                # the reference hands the rank-1 x to cKDTree.query (expert.pyx:837), which current SciPy refuses:
                # in 1D its nearest mode only works with the model index given
                from scipy.spatial import cKDTree
                I1 = cKDTree(c["x"][:, None]).query(c["xq"][:, None])[1].astype(np.int_)
                on, In = s.interpolate(c["xq"], mode="nearest", diff=d, I=I1)
            else:
                on, In = s.interpolate(c["xq"], mode="nearest", diff=d)
            assert Ic.shape == () and Ic.item() is None
            out["d%d/continuous_diff%d" % (dim, d)] = oc
            out["d%d/nearest_diff%d" % (dim, d)] = on
            out["d%d/I_nearest" % dim] = np.asarray(In, np.int64)
        out["d%d/input_checksum" % dim] = np.array([c["xk"].sum(), c["fk"].sum(), np.nansum(c["xq"]), c["r"]])
        nanq = int(np.isnan(out["d%d/continuous_diff0" % dim]).sum())
        print("dim %d: %d queries, %d without a model within r = %.3g (NaN)" % (dim, NQ, nanq, c["r"]))
        assert nanq == NFAR
    np.savez_compressed(HERE / "golden_continuous.npz", **out)
    print("written", HERE / "golden_continuous.npz")


if __name__ == "__main__":
    main()
