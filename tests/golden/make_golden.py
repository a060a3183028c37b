#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):
    python oracle/build_ref.py && python tests/golden/make_golden.py
It imports the compiled reference from oracle/_ref (oracle.load_reference()) and stores, for a set of
small seeded cases covering every dimension / order / weighting / algorithm / knowns pattern on the hot
path, the inputs and the reference's outputs:
    golden_cases.npz   per case: x, hoods, f, meta, fi0 -> fi_ref, fi_ref_perm (same fit with the neighbour
                       order permuted: the reference's own reproducibility floor), sens_ref (first 12 cases), iters_ref,
                       interpolation outputs for every derivative slot, conds (debug=True)
    golden_defs.json   every integer constant exported by wlsqm.fitter.defs
    golden_kat.npz     the README example and the lapackdrivers known answers
The GPU box has no /root/reference; tests read only these files.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "oracle")]
import oracle as orc      # noqa: E402
import workloads as wl    # noqa: E402

NSENS = 12   # sens is stored for the first NSENS cases only (fixture size)

CASES = [
    # name, dim, order, k, knowns, wm, algo, max_iter, n
    ("d2o2k12_bF_center", 2, 2, 12, 1, 2, 1, 0, 150),
    ("d2o4k30_k0_uniform", 2, 4, 30, 0, 1, 1, 0, 150),
    ("d2o4k30_k0_center", 2, 4, 30, 0, 2, 1, 0, 150),
    ("d2o4k30_bF_uniform_iter3", 2, 4, 30, 1, 1, 2, 3, 150),
    ("d2o3k24_bF_uniform", 2, 3, 24, 1, 1, 1, 0, 150),
    ("d2o3k24_bY_uniform", 2, 3, 24, 4, 1, 1, 0, 100),
    ("d2o1k8_k0_center", 2, 1, 8, 0, 2, 1, 0, 100),
    ("d2o0k6_k0_uniform", 2, 0, 6, 0, 1, 1, 0, 100),
    ("d3o4k60_bF_center_iter3", 3, 4, 60, 1, 2, 2, 3, 90),
    ("d3o4k60_k0_uniform", 3, 4, 60, 0, 1, 1, 0, 90),
    ("d3o3k40_bF_uniform", 3, 3, 40, 1, 1, 1, 0, 80),
    ("d3o2k20_bFXY_center", 3, 2, 20, 1 | (1 << 5), 2, 1, 0, 100),
    ("d1o3k8_bF_uniform", 1, 3, 8, 1, 1, 1, 0, 150),
    ("d1o4k10_k0_center_iter5", 1, 4, 10, 0, 2, 2, 5, 100),
    ("d1o2k5_bX_uniform", 1, 2, 5, 2, 1, 1, 0, 100),
]


def main():
    ref = orc.load_reference()
    if ref is None:
        raise SystemExit("oracle/_ref is not built: run python oracle/build_ref.py first")
    out = {}
    names = []
    for name, dim, order, k, knowns, wm, algo, max_iter, n in CASES:
        x = wl.cloud(n, dim, seed=42)
        hoods = wl.hoods_knn(x, k)
        f = wl.field(x)
        no = ref.number_of_dofs(dim, order)
        nk = np.full(n, k, np.int32)
        od = np.full(n, order, np.int32)
        kn = np.full(n, knowns, np.int64)
        w = np.full(n, wm, np.int32)
        rng = np.random.default_rng(1)
        fi0 = 0.1 * rng.standard_normal((n, no))
        fi0[:, 0] = f
        # prescribed knowns other than F: analytic-looking but arbitrary values are fine for parity
        xk, fk = x[hoods], f[hoods]

        def run(xk_, fk_):
            s = ref.ExpertSolver(dim, nk, od, kn, w, algorithm=algo, do_sens=True, max_iter=max_iter, ntasks=1, debug=True)
            s.prepare(x, np.ascontiguousarray(xk_))
            fi = fi0.copy()
            sens = np.zeros((n, k, no))
            it = s.solve(np.ascontiguousarray(fk_), fi, sens)
            return s, fi, sens, it

        s, fi, sens, it = run(xk, fk)
        perm = np.random.default_rng(7).permutation(k)
        _, fi_p, _, _ = run(xk[:, perm], fk[:, perm])
        conds = s.conds()
        # interpolation: 3 queries per model, every derivative slot, model index from the reference's kd-tree
        s.prep_interpolate()
        rq = np.random.default_rng(5)
        xq = np.repeat(x.reshape(n, -1), 3, axis=0) + 0.3 * wl.H0 * rq.uniform(-1, 1, (3 * n, dim))
        xq_arg = xq if dim > 1 else np.ascontiguousarray(xq[:, 0])
        size = ref.number_of_dofs(dim, 4)
        interp = np.empty((size, 3 * n))
        I = None
        if dim == 1:
            # the reference's own nearest-model search rejects 1-D query arrays (cKDTree wants (nx,1));
            # give it the index explicitly, from the same kd-tree
            I = np.ascontiguousarray(s.tree.query(xq)[1], dtype=np.int_)
        for d in range(size):
            o, I_out = s.interpolate(xq_arg, mode='nearest', diff=d, I=I)
            I = I_out
            interp[d] = o
        for key, val in dict(x=x, hoods=hoods, f=f, fi0=fi0, fi_ref=fi, fi_ref_perm=fi_p, sens_ref=sens[:NSENS],
                             iters_ref=np.int32(it), conds_ref=conds, xq=xq_arg, I_ref=np.asarray(I, np.int64),
                             interp_ref=interp,
                             meta=np.array([dim, order, k, knowns, wm, algo, max_iter, n], np.int64)).items():
            out[f"{name}/{key}"] = val
        names.append(name)
        print(f"{name}: iters {it}, cond median {np.median(conds):.3g} max {conds.max():.3g}")
    out["names"] = np.array(names)
    np.savez_compressed(HERE / "golden_cases.npz", **out)

    import wlsqm.fitter.defs as rd
    defs = {k: int(getattr(rd, k)) for k in dir(rd) if not k.startswith("_") and isinstance(getattr(rd, k), int)}
    (HERE / "golden_defs.json").write_text(json.dumps(defs, indent=0, sort_keys=True))

    # known answers outside the solver classes
    kat = {}
    # README example (README.md:102-138 style): exact quadratic recovered by fit_2D order 2
    rng = np.random.default_rng(42)
    xk = rng.uniform(-1, 1, (20, 2))
    xi = np.array([0.1, -0.2])
    coef = np.array([1.0, 2.0, 3.0, 10.0, 4.0, 12.0])     # f, fx, fy, fxx, fxy, fyy at xi
    dx, dy = xk[:, 0] - xi[0], xk[:, 1] - xi[1]
    fk = coef[0] + coef[1] * dx + coef[2] * dy + 0.5 * coef[3] * dx * dx + coef[4] * dx * dy + 0.5 * coef[5] * dy * dy
    fi = np.zeros(6)
    ref.fit_2D(xk, fk, xi, fi, None, do_sens=0, order=2, knowns=0, weighting_method=ref.WEIGHT_UNIFORM)
    kat.update(readme_xk=xk, readme_fk=fk, readme_xi=xi, readme_fi=fi, readme_expected=coef)
    # batched general drivers
    import wlsqm.utils.lapackdrivers as ld
    for nn in (3, 15, 36):
        A = np.asfortranarray(rng.standard_normal((nn, nn, 12)))
        b = np.asfortranarray(rng.standard_normal((nn, 12)))
        LU = A.copy(order='F')
        ipiv = np.zeros((nn, 12), np.int32, order='F')
        ld.mgeneralfactor(LU, ipiv)
        xs = b.copy(order='F')
        ld.mgeneralfactored(LU, ipiv, xs)
        kat.update({f"lu{nn}_A": A, f"lu{nn}_b": b, f"lu{nn}_LU": LU, f"lu{nn}_ipiv": ipiv, f"lu{nn}_x": xs})
    np.savez_compressed(HERE / "golden_kat.npz", **kat)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
