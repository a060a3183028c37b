#!/usr/bin/env python
"""Parity + conditioning report for every BASELINE.json config ("report conds alongside every parity
result").  Runs on the GPU box:   python tools/parity_report.py [--cfg3-n 4000000] > profiles/r02_parity_report.txt

Per config and derivative order d: DOF-scaled |gpu - oracle| (p50 / p99 / max), the oracle's own neighbour-permutation
self-difference measured in the same run (the floor), the ratio of the p99s, and the 2-norm condition numbers of the
scaled matrices (ExpertSolver.conds(), debug=True) of the compared cases.  The GPU runs each config at its FULL size
on real kNN neighbourhoods; the oracle (scalar C) runs on a strided subsample of the cases.
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "python-wlsqm_b200"), str(ROOT / "oracle"), str(ROOT / "tests")]
import wlsqm_b200 as wlsqm  # noqa: E402
import workloads as wl      # noqa: E402
import parity               # noqa: E402
import oracle as orc        # noqa: E402


def q3(v):
    v = np.asarray(v)
    v = v[np.isfinite(v)]
    return (np.median(v), np.quantile(v, 0.99), v.max()) if v.size else (np.nan,) * 3


def report(label, dim, order, got, ref, ref_perm, conds, extra=""):
    rep = wl.parity_report(got, ref, dim, order)
    floor = wl.parity_report(ref_perm, ref, dim, order)
    c50, c99, cmx = q3(conds)
    print("## %s" % label)
    print("   cases compared: %d   conds() of the scaled matrices p50 / p99 / max = %.3g / %.3g / %.3g   %s" % (len(got), c50, c99, cmx, extra))
    print("   d | gpu-vs-oracle p50 / p99 / max      | oracle self-noise (floor) p50 / p99 / max | p99 ratio | within 4x floor")
    for d in sorted(rep):
        g, f = rep[d], floor.get(d, (0.0, 0.0, 0.0))
        ratio = g[1] / f[1] if f[1] > 0 else float("inf") if g[1] > 0 else 0.0
        lim = (parity.FLOOR_FACTOR * f[0] + parity.ABS_FLOOR, parity.FLOOR_FACTOR * f[1] + 10 * parity.ABS_FLOOR,
               parity.FLOOR_FACTOR * f[2] + 100 * parity.ABS_FLOOR)
        ok = g[0] <= lim[0] and g[1] <= lim[1] and g[2] <= lim[2]
        print("   %d | %.2e / %.2e / %.2e | %.2e / %.2e / %.2e          | %9.2f | %s" % (d, *g, *f, ratio, "yes" if ok else "NO"))
    print(flush=True)


def conds_of(dim, meta, xi, xk):
    nk, od, kn, wm = meta
    s = wlsqm.ExpertSolver(dim, nk, od, kn, wm, debug=True)
    s.prepare(xi, xk)
    return s.conds()


def subsample_compare(label, dim, order, idx, meta, x_d, hoods_d, f_d, fi_in_d, fi_out_d, algo, max_iter, iters=None, sens_d=None):
    it = torch.from_numpy(idx).cuda()
    hd = hoods_d[it].long()
    x2 = x_d if x_d.dim() == 2 else x_d[:, None]
    xk_s = x2[hd].cpu().numpy()
    if dim == 1:
        xk_s = np.ascontiguousarray(xk_s[:, :, 0])
    fk_s, xi_s = f_d[hd].cpu().numpy(), x_d[it].cpu().numpy()
    m = tuple(a[idx] for a in meta)
    fi0 = fi_in_d[it].cpu().numpy()
    ref, sens_o, it_o, so = parity.oracle_solve(dim, *m, xi_s, xk_s, fk_s, fi0, algo, sens_d is not None, max_iter)
    a, b = parity.hetero_self_noise(dim, *m, xi_s, xk_s, fk_s, fi0, algo, max_iter)
    got = fi_out_d[it].cpu().numpy()
    cond = conds_of(dim, m, xi_s, xk_s)
    extra = ""
    if iters is not None:
        gi, oi = iters[idx], so.iters
        pairs = {}
        for g_, o_ in zip(gi.tolist(), oi.tolist()):
            pairs[(g_, o_)] = pairs.get((g_, o_), 0) + 1
        extra = ("refinement iterations: gpu == oracle for %d of %d cases, max gpu %d / oracle %d; (gpu, oracle) -> cases: %s "
                 "[the exit `norm == prev_norm` (impl.pyx:1057-1060) compares bit patterns, so it can fire one round apart]"
                 % (int((gi == oi).sum()), len(idx), int(gi.max()), int(oi.max()), sorted(pairs.items())))
    if sens_d is not None:
        sg = sens_d[it].cpu().numpy()
        same_nan = np.array_equal(np.isnan(sg), np.isnan(sens_o))
        r = np.nan_to_num(sens_o)
        sc = np.abs(r).max(axis=(0, 1)); sc[sc == 0] = 1
        e = np.abs(np.nan_to_num(sg) - r) / sc
        extra += "   sens: NaN pattern equal %s, scaled error p50 / max = %.2e / %.2e" % (same_nan, np.median(e), e.max())
    for knv in np.unique(m[2]):
        g = m[2] == knv
        report("%s, knowns=%d" % (label, int(knv)), dim, order, got[g], ref[g], (b + (ref - a))[g], cond[g], extra)


def meta(n, k, order, knowns, wm):
    return (np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, knowns, np.int64), np.full(n, wm, np.int32))


def expert_cfg(label, n, dim, order, k, knowns, wm, algo, do_sens, every, max_iter=3):
    t0 = time.time()
    x = wl.cloud(n, dim)
    x_d = torch.from_numpy(x).cuda()
    hoods_d = wlsqm.knn_hoods(x_d, k)
    m = meta(n, k, order, knowns, wm)
    f_d = torch.from_numpy(wl.field(x)).cuda()
    no = wlsqm.number_of_dofs(dim, order)
    fi_in = torch.zeros((n, no), dtype=torch.float64, device="cuda")
    fi_in[:, 0] = f_d
    fi = fi_in.clone()
    s = wlsqm.ExpertSolver(dim, *m, algorithm=algo, do_sens=do_sens, max_iter=max_iter)
    s.prepare_hoods(x_d, hoods_d)
    sens = torch.empty((n, k, no), dtype=torch.float64, device="cuda") if do_sens else None
    s.solve_hoods(f_d, fi, sens)
    its = s.iterations() if algo == wlsqm.ALGO_ITERATIVE else None
    subsample_compare("%s [GPU at n = %d, oracle on every %dth case]" % (label, n, every), dim, order, np.arange(0, n, every), m, x_d,
                      hoods_d, f_d, fi_in, fi, algo, max_iter, its, sens)
    print("   (%.1f s)\n" % (time.time() - t0), flush=True)
    del s, sens
    torch.cuda.empty_cache()
    wlsqm.pool_trim()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg3-n", type=int, default=4_000_000)
    a = ap.parse_args()
    print("# Parity report, GPU (wlsqm_b200 on %s) vs the oracle (oracle/wlsqm_oracle.c, pinned against the unmodified reference)" % torch.cuda.get_device_name(0))
    print("# criterion: tests/parity.py (each quantile <= 4 x the oracle's own neighbour-permutation self-difference + a few ulp)\n")
    # cfg1: the one-shot API on the example-sized cloud, every case compared
    n, k = 10_000, 12
    x = wl.cloud(n, 2, unit_box=True)
    hoods = wl.hoods_knn(x, k)
    f = wl.field(x)
    xk, fk = np.ascontiguousarray(x[hoods]), np.ascontiguousarray(f[hoods])
    m = meta(n, k, 2, wlsqm.b2_F, wlsqm.WEIGHT_CENTER)
    fi0 = np.zeros((n, 6)); fi0[:, 0] = f
    got = fi0.copy()
    wlsqm.fit_2D_many_parallel(xk, fk, m[0], x, got, None, 0, m[1], m[2], m[3], ntasks=8)
    ref, _, _, _ = parity.oracle_solve(2, *m, x, xk, fk, fi0)
    pa, pb = parity.permuted_self_noise(2, *m, x, xk, fk, fi0)
    report("cfg1: fit_2D_many_parallel, 10k points, order 2, k=12, b2_F, WEIGHT_CENTER [all cases]", 2, 2, got, ref, pb + (ref - pa),
           conds_of(2, m, x, xk))
    expert_cfg("cfg2 headline: 2D order 4, k=30, knowns=0, WEIGHT_UNIFORM, ALGO_BASIC", 1_000_000, 2, 4, 30, 0, 1, 1, False, 499)
    expert_cfg("cfg2 variant: knowns=b2_F", 1_000_000, 2, 4, 30, 1, 1, 1, False, 499)
    expert_cfg("cfg2 variant: WEIGHT_CENTER, ALGO_ITERATIVE(3)", 1_000_000, 2, 4, 30, 0, 2, 2, False, 499)
    free, _ = torch.cuda.mem_get_info()
    n3 = a.cfg3_n if free > 160e9 else min(a.cfg3_n, 2_000_000)
    expert_cfg("cfg3: 3D order 4, k=60, b3_F, WEIGHT_CENTER, ALGO_ITERATIVE(3), do_sens", n3, 3, 4, 60, 1, 2, 2, True, n3 // 1000)
    for dim, k in ((2, 24), (1, 8)):
        p = parity.cfg4_problem(dim, 2_000_000, k)
        fi = p["fi_in_d"].clone()
        s = wlsqm.ExpertSolver(dim, *p["meta"])
        s.prepare_hoods(p["x_d"], p["hoods_d"])
        s.solve_hoods(p["f_d"], fi)
        idx = np.sort(np.concatenate([p["interior"][:: len(p["interior"]) // 1500], p["boundary"][:: max(1, len(p["boundary"]) // 700)]]))
        subsample_compare("cfg4: %dD order 3, k=%d, WEIGHT_UNIFORM, 2M points; F known in the interior, d/d%s known on the boundary "
                          "(neighbours from the interior only)" % (dim, k, "y" if dim == 2 else "x"), dim, 3, idx, p["meta"],
                          p["x_d"], p["hoods_d"], p["f_d"], p["fi_in_d"], fi, 1, 0)
        del s, p
        torch.cuda.empty_cache()
    # cfg5: the evaluator, every derivative slot, on the cfg2 cloud
    n, k = 1_000_000, 30
    x = wl.cloud(n, 2)
    x_d = torch.from_numpy(x).cuda()
    hoods_d = wlsqm.knn_hoods(x_d, k)
    m = meta(n, k, 4, 0, 1)
    s = wlsqm.ExpertSolver(2, *m)
    s.prepare_hoods(x_d, hoods_d)
    fi = torch.zeros((n, 15), dtype=torch.float64, device="cuda")
    s.solve_hoods(torch.from_numpy(wl.field(x)).cuda(), fi)
    I = torch.arange(n, device="cuda", dtype=torch.int64).repeat_interleave(16)
    g = torch.Generator(device="cuda").manual_seed(5)
    xq = x_d[I] + 0.3 * wl.H0 * (2 * torch.rand((16 * n, 2), dtype=torch.float64, device="cuda", generator=g) - 1)
    s.tree = object()
    allv, _ = s.interpolate(xq, diff="all", I=I)
    pick = torch.arange(0, 16 * n, 1601, device="cuda")
    so = orc.OracleSolver(2, m[0][:1], m[1][:1], m[2][:1], m[3][:1])
    so.order = m[1]
    so.xi, so.fi = x, fi.cpu().numpy()
    xq_h, I_h = xq[pick].cpu().numpy(), I[pick].cpu().numpy()
    print("## cfg5: interpolate, 16M queries in the cfg2 cloud, all 15 derivative slots in one pass vs the oracle's evaluator")
    print("   (same coefficients on both sides; %d queries compared; error relative to the largest value of the slot)" % len(pick))
    worst = 0.0
    for d in range(15):
        oo = so.interpolate(xq_h, I_h, d)
        one, _ = s.interpolate(xq[pick], diff=d, I=I[pick])
        e_all = np.abs(allv[pick, d].cpu().numpy() - oo).max() / np.abs(oo).max()
        e_one = np.abs(one.cpu().numpy() - oo).max() / np.abs(oo).max()
        worst = max(worst, e_all, e_one)
        print("   slot %2d: all-slots pass %.2e   single-slot call %.2e" % (d, e_all, e_one))
    print("   worst: %.2e (a few ulp)\n" % worst)


if __name__ == "__main__":
    main()
