"""heterogeneous batch (per-case nk / order / knowns / weighting vary): prepare and solve throughput, 1M cases 2D"""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/python-wlsqm_b200"]
import numpy as np, torch, wlsqm_b200 as wlsqm
n, kmax = 1_000_000, 30
rng = np.random.default_rng(0)
g = torch.Generator(device="cuda").manual_seed(0)
xi = 10 * torch.rand((n, 2), dtype=torch.float64, device="cuda", generator=g)
xk = xi[:, None, :] + 0.015 * (2 * torch.rand((n, kmax, 2), dtype=torch.float64, device="cuda", generator=g) - 1)
fk = torch.sin(xk[..., 0]) * torch.cos(xk[..., 1])
od = rng.integers(2, 5, n).astype(np.int32)
nk = rng.integers(22, kmax + 1, n).astype(np.int32)
kn = rng.integers(0, 2, n).astype(np.int64)            # F known or not
wm = rng.integers(1, 3, n).astype(np.int32)
fi = torch.zeros((n, 15), dtype=torch.float64, device="cuda")
s = wlsqm.ExpertSolver(2, nk, od, kn, wm)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
tp = t(lambda: s.prepare(xi, xk), 3)
tsv = t(lambda: s.solve(fk, fi))
no = np.array([1, 3, 6, 10, 15])[od]; nr = no - kn
byt = float((8 * (nr * nk + nr * kn + nk + 2 * no)).sum())
print("heterogeneous 1M cases (order 2-4, nk 22-30, F known or not, both weightings): prepare %.2f ms (%.3g fits/s), solve %.3f ms "
      "(%.3g points/s, %.0f GB/s algorithmic = %.2f of the HBM peak)" % (tp, n / tp * 1e3, tsv, n / tsv * 1e3, byt / tsv / 1e6, byt / tsv / 1e6 / 6526))
# launch-shape sweep of the general (per-case record) solve kernel through the library's A/B switches
import os, itertools
if len(sys.argv) > 1 and sys.argv[1] == "sweep":
    for S, W, M in itertools.product((2, 3, 4), (8, 12, 16, 24), (16, 24, 32, 48)):
        if M < W:
            continue
        os.environ.update(WLSQM_SOLVE_STAGES=str(S), WLSQM_SOLVE_WARPS=str(W), WLSQM_SOLVE_MAXWARPS_SM=str(M))
        try:
            ts_ = t(lambda: s.solve(fk, fi), 3)
        except Exception as e:          # noqa: BLE001
            print("stages %d warps %d max warps/SM %d: %s" % (S, W, M, e)); continue
        print("stages %d warps/CTA %d max warps/SM %d: solve %.3f ms (%.2f of the HBM peak)" % (S, W, M, ts_, byt / ts_ / 1e6 / 6526))
