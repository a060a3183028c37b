"""Where does a one-shot fit_2D_many_parallel call (cfg1: 10k points, order 2, k=12) spend its time?  GPU box only."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "python-wlsqm_b200")]
import torch
import wlsqm_b200 as wlsqm
import workloads as wl

with_torch_pool = "--torchpool" in sys.argv
if with_torch_pool:
    big = torch.empty(8 << 30, dtype=torch.uint8, device="cuda"); del big   # torch caching allocator holds 8 GB
torch.cuda.synchronize()
for n in (10_000, 100_000):
    k = 12
    x = wl.cloud(n, 2, unit_box=True); h = wl.hoods_knn(x, k); f = wl.field(x)
    xk, fk = np.ascontiguousarray(x[h]), np.ascontiguousarray(f[h])
    fi = np.zeros((n, 6)); fi[:, 0] = f
    meta = (np.full(n, k, np.int32), np.full(n, 2, np.int32), np.full(n, wlsqm.b2_F, np.int64), np.full(n, wlsqm.WEIGHT_CENTER, np.int32))
    def T(fn, reps=7):
        fn(); fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        return 1e3 * float(np.median(ts))
    print(n, "fit_many host arrays  %.3f ms" % T(lambda: wlsqm.fit_2D_many_parallel(xk, fk, meta[0], x, fi, None, 0, meta[1], meta[2], meta[3])))
    xk_d, fk_d, x_d, fi_d = (torch.from_numpy(a).cuda() for a in (xk, fk, x, fi))
    print(n, "fit_many cuda tensors %.3f ms" % T(lambda: wlsqm.fit_2D_many_parallel(xk_d, fk_d, meta[0], x_d, fi_d, None, 0, meta[1], meta[2], meta[3])))
    # phases through ExpertSolver
    box = {}
    def create(): box["s"] = wlsqm.ExpertSolver(2, *meta, algorithm=wlsqm.ALGO_BASIC, do_sens=False)
    print(n, "  create+destroy      %.3f ms" % T(create))
    s = box["s"]
    print(n, "  prepare (host)      %.3f ms" % T(lambda: s.prepare(x, xk)))
    print(n, "  solve   (host)      %.3f ms" % T(lambda: s.solve(fk, fi)))
    print(n, "  prepare (cuda)      %.3f ms" % T(lambda: s.prepare(x_d, xk_d)))
    print(n, "  solve   (cuda)      %.3f ms" % T(lambda: s.solve(fk_d, fi_d)))
