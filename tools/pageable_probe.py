"""solve() through ordinary (pageable) numpy arrays vs pinned ones, 1M points, 2D order 4, k = 30"""
import sys, time
sys.path[:0] = ["/root/repo", "/root/repo/python-wlsqm_b200"]
import numpy as np, torch, wlsqm_b200 as wlsqm
n, k = 1_000_000, 30
rng = np.random.default_rng(0)
xi = 10 * rng.random((n, 2)); xk = xi[:, None, :] + 0.015 * (2 * rng.random((n, k, 2)) - 1)
fk = np.sin(xk[..., 0]) * np.cos(xk[..., 1]); fi = np.zeros((n, 15))
m = (np.full(n, k, np.int32), np.full(n, 4, np.int32), np.zeros(n, np.int64), np.full(n, 1, np.int32))
s = wlsqm.ExpertSolver(2, *m)
t0 = time.perf_counter(); s.prepare(xi, xk); print("prepare (pageable xk 480 MB): %.1f ms" % (1e3 * (time.perf_counter() - t0)))
for name, a, b in (("pageable", fk, fi), ("pinned", None, None)):
    if a is None:
        a = wlsqm.pinned_empty((n, k)); a[...] = fk
        b = wlsqm.pinned_empty((n, 15)); b[...] = 0
    s.solve(a, b)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); s.solve(a, b); ts.append(time.perf_counter() - t0)
    print("solve %s: %.2f ms/step -> %.3g points/s" % (name, 1e3 * min(ts), n / min(ts)))
