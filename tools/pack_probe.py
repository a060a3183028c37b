"""launch-shape sweep of the packed solve (cfg4): WLSQM_SOLVE_STAGES x WLSQM_SOLVE_WARPS, 2M points, 1D order 3 k=8 and 2D order 3 k=24"""
import os, sys, itertools
sys.path[:0] = ["/root/repo", "/root/repo/python-wlsqm_b200"]
import numpy as np, torch, wlsqm_b200 as wlsqm

def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

def run(dim, order, k, n=2_000_000):
    g = torch.Generator(device="cuda").manual_seed(0)
    xi = 10 * torch.rand((n, dim), dtype=torch.float64, device="cuda", generator=g)
    xk = xi[:, None, :] + 0.015 * (2 * torch.rand((n, k, dim), dtype=torch.float64, device="cuda", generator=g) - 1)
    fk = torch.sin(xk[..., 0])
    if dim == 1:
        xi, xk = xi[:, 0].contiguous(), xk[..., 0].contiguous()
    no = wlsqm.number_of_dofs(dim, order)
    kn = np.ones(n, np.int64); kn[::1000] = 2
    nk, od, wm = np.full(n, k, np.int32), np.full(n, order, np.int32), np.ones(n, np.int32)
    fi = torch.zeros((n, no), dtype=torch.float64, device="cuda")
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    s.prepare(xi, xk)
    nr = no - 1
    byt = 8.0 * (nr * k + nr + k + 2 * no) * n
    for key in ("WLSQM_SOLVE_STAGES", "WLSQM_SOLVE_WARPS", "WLSQM_SOLVE_MAXWARPS_SM"):
        os.environ.pop(key, None)
    ts_ = t(lambda: s.solve(fk, fi))
    print("%dD o%d k%d default: %.4f ms (%.2f of 6526 GB/s)" % (dim, order, k, ts_, byt / ts_ / 1e6 / 6526), flush=True)
    for S, W, M in itertools.product((2, 3, 4, 6), (8, 16, 32), (32, 64)):
        os.environ.update(WLSQM_SOLVE_STAGES=str(S), WLSQM_SOLVE_WARPS=str(W), WLSQM_SOLVE_MAXWARPS_SM=str(M))
        try:
            ts_ = t(lambda: s.solve(fk, fi))
        except Exception as e:      # noqa: BLE001
            print("  S %d W %d M %d: %s" % (S, W, M, e)); continue
        print("  stages %d warps/CTA %d max warps/SM %d: %.4f ms (%.2f)" % (S, W, M, ts_, byt / ts_ / 1e6 / 6526), flush=True)

run(1, 3, 8)
run(2, 3, 24)
