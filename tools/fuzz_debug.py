"""diagnostics for tests/test_gpu_parity.py::test_random_knowns_masks_orders_and_sizes: the worst cases of a seed, per kernel"""
import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/tests", "/root/repo/oracle", "/root/repo/python-wlsqm_b200"]
import numpy as np
import parity
import oracle as orc
import wlsqm_b200 as wlsqm


def build(dim, seed):
    n = 1500 if dim == 3 else 2500
    kmax = {1: 12, 2: 30, 3: 60}[dim]
    nomax = orc.number_of_dofs(dim, 4)
    x, hoods, f = parity.make_case(n, dim, kmax, seed=seed)
    xk, fk = parity.gathered(x, f, hoods)
    rng = np.random.default_rng(seed)
    od = rng.integers(0, 5, n).astype(np.int32)
    no = np.array([orc.number_of_dofs(dim, int(o)) for o in od])
    kn = np.zeros(n, np.int64)
    for j in range(n):
        bits = rng.random(no[j]) < rng.choice([0.0, 0.15, 0.5, 0.9])
        if bits.all():
            bits[rng.integers(no[j])] = False
        kn[j] = int(sum(1 << o for o in range(no[j]) if bits[o]))
    nr = np.array([no[j] - bin(int(kn[j])).count("1") for j in range(n)])
    lo = np.minimum(kmax, np.maximum(nr + 2, (3 * nr) // 2 + 1))
    nk = rng.integers(lo, kmax + 1).astype(np.int32)
    wm = rng.integers(1, 3, n).astype(np.int32)
    fi0 = rng.standard_normal((n, nomax))
    exact, _, _, _ = parity.oracle_solve(dim, np.full(n, kmax, np.int32), np.full(n, 4, np.int32), np.zeros(n, np.int64),
                                         np.ones(n, np.int32), x, xk, fk, np.zeros((n, nomax)), 1)
    for j in range(n):
        for o in range(no[j]):
            if kn[j] >> o & 1:
                fi0[j, o] = exact[j, o]
    return dict(dim=dim, n=n, x=x, xk=xk, fk=fk, od=od, no=no, kn=kn, nr=nr, nk=nk, wm=wm, fi0=fi0)


def gpu(c, algo, sub=None):
    sel = slice(None) if sub is None else sub
    s = wlsqm.ExpertSolver(c["dim"], c["nk"][sel], c["od"][sel], c["kn"][sel], c["wm"][sel], algorithm=algo, max_iter=6, ntasks=1)
    s.prepare(c["x"][sel], c["xk"][sel])
    fi = c["fi0"][sel].copy()
    s.solve(c["fk"][sel], fi)
    return fi


dim, algo, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
c = build(dim, seed)
fi_o, _, _, _ = parity.oracle_solve(dim, c["nk"], c["od"], c["kn"], c["wm"], c["x"], c["xk"], c["fk"], c["fi0"], algo, False, max_iter=6)
for kern in ("reg", "smem"):
    if kern == "smem":
        os.environ["WLSQM_PREP_KERNEL"] = "smem"
    else:
        os.environ.pop("WLSQM_PREP_KERNEL", None)
    fi_g = gpu(c, algo)
    err = np.abs(fi_g - fi_o)
    rel = err.max(axis=1) / np.maximum(np.abs(fi_o).max(axis=1), 1e-300)
    worst = np.argsort(-rel)[:6]
    print("== kernel %s: %dD algo %d seed %d" % (kern, dim, algo, seed))
    for j in worst:
        # the same case alone / among its own order only
        alone = gpu(c, algo, slice(j, j + 1))[0]
        print("case %d order %d nk %d nr %d wm %d knowns %s: max|gpu-oracle| %.3e (max|fi| %.3e); alone: %.3e; conds? slots wrong: %s" % (
            j, c["od"][j], c["nk"][j], c["nr"][j], c["wm"][j], bin(int(c["kn"][j])), err[j].max(), np.abs(fi_o[j]).max(),
            np.abs(alone - fi_o[j]).max(), np.where(err[j] > 1e-6 * np.abs(fi_o[j]).max())[0].tolist()))
