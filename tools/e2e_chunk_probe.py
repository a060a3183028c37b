"""solve() through page-locked host arrays, 1M points 2D order 4 k = 30: ms per step for the current WLSQM_SOLVE_CHUNK"""
import sys, time, os
sys.path[:0] = ["/root/repo", "/root/repo/python-wlsqm_b200"]
import numpy as np, torch, wlsqm_b200 as wlsqm
n, k = 1_000_000, 30
g = torch.Generator(device="cuda").manual_seed(0)
xi = 10 * torch.rand((n, 2), dtype=torch.float64, device="cuda", generator=g)
xk = xi[:, None, :] + 0.015 * (2 * torch.rand((n, k, 2), dtype=torch.float64, device="cuda", generator=g) - 1)
m = (np.full(n, k, np.int32), np.full(n, 4, np.int32), np.zeros(n, np.int64), np.full(n, 1, np.int32))
s = wlsqm.ExpertSolver(2, *m)
s.prepare(xi, xk)
fk = wlsqm.pinned_empty((n, k)); fk[...] = (torch.sin(xk[..., 0]) * torch.cos(xk[..., 1])).cpu().numpy()
fi = wlsqm.pinned_empty((n, 15)); fi[...] = 0
for _ in range(3): s.solve(fk, fi)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); s.solve(fk, fi); ts.append(time.perf_counter() - t0)
print("chunk %s: min %.3f ms, median %.3f ms" % (os.environ.get("WLSQM_SOLVE_CHUNK", "default"), 1e3 * min(ts), 1e3 * float(np.median(ts))))
