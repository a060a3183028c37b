"""prepare() on 400k 3D cases of mixed orders 2-4 (nk 40-60): per-order launches (default) vs one launch (WLSQM_PREP_BUCKETS=0)"""
import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/python-wlsqm_b200"]
import numpy as np, torch, wlsqm_b200 as wlsqm
n, kmax, dim = 400_000, 60, 3
rng = np.random.default_rng(0)
g = torch.Generator(device="cuda").manual_seed(0)
xi = 10 * torch.rand((n, dim), dtype=torch.float64, device="cuda", generator=g)
xk = xi[:, None, :] + 0.015 * (2 * torch.rand((n, kmax, dim), dtype=torch.float64, device="cuda", generator=g) - 1)
od = rng.integers(2, 5, n).astype(np.int32)
nk = rng.integers(40, kmax + 1, n).astype(np.int32)
kn = rng.integers(0, 2, n).astype(np.int64)
wm = rng.integers(1, 3, n).astype(np.int32)
s = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
a = t(lambda: s.prepare(xi, xk))
os.environ["WLSQM_PREP_BUCKETS"] = "0"
b = t(lambda: s.prepare(xi, xk))
print("3D mixed orders 2-4, %d cases: prepare %.2f ms with one launch per order, %.2f ms with the order-4 kernel for every case" % (n, a, b))
