"""heterogeneous batches beyond 2D ALGO_BASIC: launch-shape sweep of the per-case-record solve kernel
(3D BASIC, 2D / 3D ALGO_ITERATIVE) through the library's A/B switches; usage: python tools/hetero_probe2.py [sweep]"""
import os, sys, itertools
sys.path[:0] = ["/root/repo", "/root/repo/python-wlsqm_b200"]
import numpy as np, torch, wlsqm_b200 as wlsqm

NO = {2: np.array([1, 3, 6, 10, 15]), 3: np.array([1, 4, 10, 20, 35])}

def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

def run(dim, n, kmin, kmax, algo):
    rng = np.random.default_rng(0)
    g = torch.Generator(device="cuda").manual_seed(0)
    xi = 10 * torch.rand((n, dim), dtype=torch.float64, device="cuda", generator=g)
    xk = xi[:, None, :] + 0.015 * (2 * torch.rand((n, kmax, dim), dtype=torch.float64, device="cuda", generator=g) - 1)
    fk = torch.sin(xk[..., 0]) * torch.cos(xk[..., 1])
    od = rng.integers(2, 5, n).astype(np.int32)
    nk = rng.integers(kmin, kmax + 1, n).astype(np.int32)
    kn = rng.integers(0, 2, n).astype(np.int64)
    wm = rng.integers(1, 3, n).astype(np.int32)
    fi = torch.zeros((n, int(NO[dim][4])), dtype=torch.float64, device="cuda")
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, algorithm=algo, max_iter=3)
    s.prepare(xi, xk)
    no = NO[dim][od]; nr = no - kn
    byt = float((8 * (nr * nk + nr * kn + nk + 2 * no + (nk * dim + dim if algo == wlsqm.ALGO_ITERATIVE else 0))).sum())
    name = "%dD %s n=%d nk %d-%d" % (dim, "ITERATIVE(3)" if algo == wlsqm.ALGO_ITERATIVE else "BASIC", n, kmin, kmax)
    for k in ("WLSQM_SOLVE_STAGES", "WLSQM_SOLVE_WARPS", "WLSQM_SOLVE_MAXWARPS_SM"):
        os.environ.pop(k, None)
    ts_ = t(lambda: s.solve(fk, fi))
    print("%s default: solve %.3f ms (%.2f of the HBM peak)" % (name, ts_, byt / ts_ / 1e6 / 6526), flush=True)
    for S, W in (itertools.product((1, 2), (8, 12, 16, 24, 32)) if "sweep" in sys.argv else ()):
        os.environ.update(WLSQM_SOLVE_STAGES=str(S), WLSQM_SOLVE_WARPS=str(W), WLSQM_SOLVE_MAXWARPS_SM="32")
        try:
            ts_ = t(lambda: s.solve(fk, fi))
        except Exception as e:          # noqa: BLE001
            print("  stages %d warps %d: %s" % (S, W, e)); continue
        print("  stages %d warps/CTA %d (max 32 warps/SM): %.3f ms (%.2f)" % (S, W, ts_, byt / ts_ / 1e6 / 6526), flush=True)
    del s

run(3, 400_000, 40, 60, wlsqm.ALGO_BASIC)
run(2, 1_000_000, 22, 30, wlsqm.ALGO_ITERATIVE)
run(3, 400_000, 40, 60, wlsqm.ALGO_ITERATIVE)
