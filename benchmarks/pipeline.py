#!/usr/bin/env python
"""End-to-end pipeline on one GPU: neighbourhood construction + prepare + time steps, with the
device-side search and gathers, next to the host-side steps the reference's callers use (cKDTree + numpy gathers).

    python benchmarks/pipeline.py [--n 1000000] [--dim 2] [--order 4] [--k 30] [--steps 20] [--host-tree]

Prints one JSON line per stage: {"stage", "n", "ms", "per_s", ...}.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "python-wlsqm_b200")]
import wlsqm_b200 as wlsqm  # noqa: E402
import workloads as wl      # noqa: E402


def emit(stage, n, ms, **kw):
    print(json.dumps(dict(stage=stage, n=n, ms=round(ms, 3), per_s=n / (ms * 1e-3), **kw)), flush=True)


def ev_time(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--k", type=int, default=30)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--host-tree", action="store_true", help="also time scipy.spatial.cKDTree on the host cores")
    a = ap.parse_args()
    n, dim, order, k = a.n, a.dim, a.order, a.k
    no = wlsqm.number_of_dofs(dim, order)
    x = wl.cloud(n, dim)
    f = wl.field(x)
    xt = torch.from_numpy(x).cuda()

    # ---- neighbourhood construction ------------------------------------------------------------------------
    ms, hoods = ev_time(lambda: wlsqm.knn_hoods(xt, k))
    emit("knn_hoods (device grid, build + query)", n, ms, k=k)
    if a.host_tree:
        from scipy.spatial import cKDTree
        t0 = time.perf_counter()
        ref = cKDTree(x.reshape(n, -1)).query(x.reshape(n, -1), k + 1, workers=-1)[1][:, 1:]
        emit("cKDTree build + query (host, all cores)", n, (time.perf_counter() - t0) * 1e3, k=k,
             identical=bool(np.array_equal(ref, hoods.cpu().numpy())))

    nk, od, kn, wm = (np.full(n, k, np.int32), np.full(n, order, np.int32), np.zeros(n, np.int64), np.full(n, 1, np.int32))
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm)
    ms, _ = ev_time(lambda: s.prepare_hoods(xt, hoods))
    emit("prepare_hoods (gather x[hoods] + prepare)", n, ms)

    # ---- time steps, host data: one value per point in, the fitted coefficients out -----------------------------
    f_h = [wlsqm.pinned_empty((n,)) for _ in range(2)]
    for t in range(2):
        f_h[t][...] = wl.field_step(f, t)
    fi_h = wlsqm.pinned_empty((n, no))
    fi_h[...] = 0.0
    for t in range(3):
        s.solve_hoods(f_h[t % 2], fi_h)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(a.steps):
        s.solve_hoods(f_h[t % 2], fi_h)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / a.steps
    emit("solve_hoods step, host f -> host fi (H2D 8 B/point, D2H %d B/point)" % (8 * no), n, ms)

    # same through the reference API: the caller gathers fk = f[hoods] on the host and passes (n, k)
    hoods_h = hoods.cpu().numpy()
    t0 = time.perf_counter()
    fk_h = f_h[0][hoods_h]
    gather_ms = (time.perf_counter() - t0) * 1e3
    fk_p = wlsqm.pinned_empty((n, k))
    fk_p[...] = fk_h
    s.solve(fk_p, fi_h)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(a.steps):
        s.solve(fk_p, fi_h)
    torch.cuda.synchronize()
    ms2 = (time.perf_counter() - t0) * 1e3 / a.steps
    emit("solve step, host fk (n,k) -> host fi (reference API)", n, ms2, host_gather_ms_not_included=round(gather_ms, 2))

    # ---- device-resident steps ---------------------------------------------------------------------------------
    ft = torch.from_numpy(f).cuda()
    fit = torch.zeros((n, no), dtype=torch.float64, device="cuda")
    for t in range(3):
        s.solve_hoods(ft, fit)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(a.steps):
        s.solve_hoods(ft, fit)
    e1.record()
    torch.cuda.synchronize()
    emit("solve_hoods step, device f -> device fi (gather + solve kernels)", n, e0.elapsed_time(e1) / a.steps)


if __name__ == "__main__":
    main()
