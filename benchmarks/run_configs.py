#!/usr/bin/env python
"""Secondary configurations of BASELINE.json (configs[0], [2], [3], [4]) on one GPU: throughput and roofline
fraction per stage.  Not the driver's bench line (that is bench.py / configs[1]); this is the measurement
behind DESIGN.md section 7 for the other kernels' variants.

    python benchmarks/run_configs.py [--scale 1.0] [--only cfg3] [--no-cpu]

Beside every GPU figure stands the reference's own CPU figure for the same stage (lines with "impl": "reference-cpu"):
the unmodified reference (oracle/_ref, OpenMP) on this box's host cores, on a bounded subsample of the same workload,
chunked where one reference solver cannot hold the cases, best of ntasks in {8, all cores}.

Inputs are generated on the device: xk = xi + h*U(-1,1)^dim (same shapes and conditioning class as kNN hoods,
without a 4M-point kd-tree build on the host); data = the analytic field of workloads.py at xk.
Every line: {"config", "stage", "n", "ms", "per_s", "bytes_per_unit", "GBps", "hbm_frac"}.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "python-wlsqm_b200")]
import wlsqm_b200 as wlsqm  # noqa: E402

PEAK = 6457.7
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
NO = {1: (1, 2, 3, 4, 5), 2: (1, 3, 6, 10, 15), 3: (1, 4, 10, 20, 35)}


def field(x):
    f = torch.sin(np.pi * x[..., 0])
    if x.shape[-1] >= 2:
        f = f * torch.cos(np.pi * x[..., 1])
    if x.shape[-1] >= 3:
        f = f * torch.exp(x[..., 2])
    return f


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def emit(config, stage, n, ms, bytes_per_unit, extra=None):
    gbps = bytes_per_unit * n / (ms * 1e-3) / 1e9 if bytes_per_unit else None
    line = {"config": config, "stage": stage, "n": n, "ms": round(ms, 4), "per_s": n / (ms * 1e-3),
            "bytes_per_unit": bytes_per_unit, "GBps": gbps, "hbm_frac": (gbps / PEAK if gbps else None)}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


# ---- the reference on the host cores (the CPU figure beside every GPU figure) -------------------------------------------
_REF = [None, False]


def reference():
    if not _REF[1]:
        _REF[1] = True
        try:
            sys.path.insert(0, str(ROOT / "oracle"))
            import os
            os.environ.setdefault("OMP_WAIT_POLICY", "passive")
            os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
            import oracle as orc
            _REF[0] = orc.load_reference()
        except Exception as exc:
            print("reference unavailable: %r" % (exc,), file=sys.stderr)
    return _REF[0]


def cpu_emit(config, stage, n, sec, extra=None):
    line = {"impl": "reference-cpu", "config": config, "stage": stage, "n": n, "ms": round(1e3 * sec, 3), "per_s": n / sec}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def cpu_expert(config, n, dim, order, k, knowns, wm, algo, do_sens, max_iter=3, mixed=None, note=""):
    """prepare + solve of the reference on n cases (one ExpertSolver: n must fit its int-sized arena)"""
    import os
    import time
    ref = reference()
    if ref is None:
        return
    g = torch.Generator().manual_seed(0)
    h = 1e-2
    xi = (h * n ** (1.0 / dim) * torch.rand((n, dim), dtype=torch.float64, generator=g))
    xk = xi[:, None, :] + 1.5 * h * (2 * torch.rand((n, k, dim), dtype=torch.float64, generator=g) - 1)
    fk = field(xk).numpy().copy()
    no = NO[dim][order]
    fi = np.zeros((n, no))
    fi[:, 0] = field(xi).numpy()
    xi_h, xk_h = xi.numpy(), xk.numpy()
    if dim == 1:
        xi_h, xk_h = np.ascontiguousarray(xi_h[:, 0]), np.ascontiguousarray(xk_h[:, :, 0])
    nk, od, kn, w = np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, knowns, np.int64), np.full(n, wm, np.int32)
    if mixed is not None:
        kn[::mixed[0]] = mixed[1]
    sens = np.zeros((n, k, no)) if do_sens else None
    cores = os.cpu_count() or 1
    best = None
    for nt in sorted({min(8, cores), cores}):
        s = ref.ExpertSolver(dim, nk, od, kn, w, algorithm=algo, do_sens=do_sens, max_iter=max_iter, ntasks=nt)
        t0 = time.perf_counter()
        s.prepare(xi_h, xk_h)
        tp = time.perf_counter() - t0
        s.solve(fk, fi, sens)
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            s.solve(fk, fi, sens)
            ts.append(time.perf_counter() - t0)
        del s
        if best is None or min(ts) < best[1]:
            best = (tp, min(ts), nt)
    tp, tsv, nt = best
    ex = {"ntasks": nt, "host_cores": cores, "sample": "%d cases in one reference ExpertSolver%s" % (n, note)}
    cpu_emit(config, "prepare", n, tp, ex)
    cpu_emit(config, "solve", n, tsv, ex)


def make(n, dim, k, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    h = 1e-2
    scale = h * n ** (1.0 / dim)
    xi = scale * torch.rand((n, dim), dtype=torch.float64, device="cuda", generator=g)
    xk = xi[:, None, :] + 1.5 * h * (2 * torch.rand((n, k, dim), dtype=torch.float64, device="cuda", generator=g) - 1)
    return xi, xk


def run_expert(config, n, dim, order, k, knowns, wm, algo, do_sens, max_iter=3, mixed=None):
    no = NO[dim][order]
    xi, xk = make(n, dim, k)
    if dim == 1:
        xi1, xk1 = xi[:, 0].contiguous(), xk[:, :, 0].contiguous()
    nk = np.full(n, k, np.int32)
    od = np.full(n, order, np.int32)
    kn = np.full(n, knowns, np.int64)
    if mixed is not None:
        kn[::mixed[0]] = mixed[1]
    w = np.full(n, wm, np.int32)
    nkn = bin(knowns).count("1")
    nr = no - nkn
    s = wlsqm.ExpertSolver(dim, nk, od, kn, w, algorithm=algo, do_sens=do_sens, max_iter=max_iter)
    args = (xi1, xk1) if dim == 1 else (xi, xk)
    ms = timeit(lambda: s.prepare(*args), reps=3, warm=1)
    emit(config, "prepare", n, ms, None, {"fits_per_s": n / (ms * 1e-3)})
    fk = field(xk).contiguous()
    fi = torch.zeros((n, no), dtype=torch.float64, device="cuda")
    fi[:, 0] = field(xi)
    sens = torch.empty((n, k, no), dtype=torch.float64, device="cuda") if do_sens else None
    b = 8 * (nr * k + nr * nkn + k + 2 * no)
    if algo == wlsqm.ALGO_ITERATIVE:
        b += 8 * (k * dim + dim)
    if do_sens:
        b += 8 * k * no
    it = [0]

    def step():
        it[0] = s.solve(fk, fi, sens)
    ms = timeit(step)
    emit(config, "solve", n, ms, b, {"iterations": it[0], "nr": nr, "algorithm": algo, "do_sens": bool(do_sens),
                                     "memory_used": s.memory_used()})
    return s, xi, fi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-cpu", action="store_true", help="skip the reference's CPU figures")
    a = ap.parse_args()
    cpu = not a.no_cpu
    sc = a.scale
    want = lambda c: (not a.only) or a.only in c

    if want("cfg1"):   # fit_2D_many_parallel, 10k points, order 2, k=12, b2_F, CENTER  (one-shot API, host arrays)
        n, k = 10_000, 12
        xi, xk = make(n, 2, k)
        xk_h, xi_h = xk.cpu().numpy(), xi.cpu().numpy()
        fk_h = field(xk).cpu().numpy()
        fi_h = np.zeros((n, 6))
        fi_h[:, 0] = field(xi).cpu().numpy()
        meta = (np.full(n, k, np.int32), np.full(n, 2, np.int32), np.full(n, wlsqm.b2_F, np.int64),
                np.full(n, wlsqm.WEIGHT_CENTER, np.int32))
        f = lambda: wlsqm.fit_2D_many_parallel(xk_h, fk_h, meta[0], xi_h, fi_h, None, 0, meta[1], meta[2], meta[3], ntasks=8)
        emit("cfg1 fit_2D_many_parallel 10k o2 k12 (host numpy arrays, one-shot)", "fit", n, timeit(f), None)
        fi_d = torch.from_numpy(fi_h).cuda()
        fk_d = torch.from_numpy(fk_h).cuda()
        f = lambda: wlsqm.fit_2D_many_parallel(xk, fk_d, meta[0], xi, fi_d, None, 0, meta[1], meta[2], meta[3], ntasks=8)
        emit("cfg1 fit_2D_many_parallel 10k o2 k12 (CUDA tensors, one-shot)", "fit", n, timeit(f), None)
        if cpu and reference() is not None:
            import os
            import time
            ref = reference()
            best = None
            for nt in sorted({1, 8, os.cpu_count() or 1}):
                for _ in range(3):
                    t0 = time.perf_counter()
                    ref.fit_2D_many_parallel(xk_h, fk_h, meta[0], xi_h, fi_h, None, 0, meta[1], meta[2], meta[3], ntasks=nt)
                    d = time.perf_counter() - t0
                    if best is None or d < best[0]:
                        best = (d, nt)
            cpu_emit("cfg1 fit_2D_many_parallel 10k o2 k12", "fit", n, best[0], {"ntasks": best[1], "sample": "the whole configuration"})
    if want("cfg2v"):
        run_expert("cfg2 variant 2D o4 k30 b2_F UNIFORM BASIC", int(1_000_000 * sc), 2, 4, 30, 1, 1, 1, False)
        run_expert("cfg2 2D o4 k30 knowns=0 CENTER ITERATIVE(3)", int(1_000_000 * sc), 2, 4, 30, 0, 2, 2, False)
        if cpu:
            cpu_expert("cfg2 variant 2D o4 k30 b2_F UNIFORM BASIC", 100_000, 2, 4, 30, 1, 1, 1, False, note=" (of 1M; <= 300k per solver)")
            cpu_expert("cfg2 2D o4 k30 knowns=0 CENTER ITERATIVE(3)", 100_000, 2, 4, 30, 0, 2, 2, False, note=" (of 1M)")
    if want("cfg3"):
        n3 = int(2_000_000 * sc)     # --scale 2.0 = the full 4M-point configuration (140 GB: fits one B200)
        run_expert("cfg3 3D o4 k60 b3_F ITERATIVE(3) do_sens (%.3gM of 4M points on one GPU)" % (n3 / 1e6), n3, 3, 4, 60,
                   1, 2, 2, True)
        torch.cuda.empty_cache()
        run_expert("cfg3-basic 3D o4 k60 b3_F BASIC no sens", int(2_000_000 * sc), 3, 4, 60, 1, 2, 1, False)
        if cpu:
            # one reference solver holds at most ~76k cases of this configuration (27 928 B per case in an int-sized arena)
            cpu_expert("cfg3 3D o4 k60 b3_F ITERATIVE(3) do_sens", 65_000, 3, 4, 60, 1, 2, 2, True,
                       note=" = one chunk of the 62 the 4M points need; scale linearly")
            cpu_expert("cfg3-basic 3D o4 k60 b3_F BASIC no sens", 65_000, 3, 4, 60, 1, 2, 1, False, note=" = one chunk")
    if want("cfg4"):
        run_expert("cfg4 2D o3 k24 UNIFORM, b2_F interior + b2_Y every 1000th", int(2_000_000 * sc), 2, 3, 24, 1, 1, 1, False,
                   mixed=(1000, wlsqm.b2_Y))
        run_expert("cfg4 1D o3 k8 UNIFORM, b1_F interior + b1_X every 1000th", int(2_000_000 * sc), 1, 3, 8, 1, 1, 1, False,
                   mixed=(1000, wlsqm.b1_X))
        if cpu:
            cpu_expert("cfg4 2D o3 k24 UNIFORM, b2_F interior + b2_Y every 1000th", 200_000, 2, 3, 24, 1, 1, 1, False,
                       mixed=(1000, wlsqm.b2_Y), note=" (of 2M; <= 600k per solver)")
            cpu_expert("cfg4 1D o3 k8 UNIFORM, b1_F interior + b1_X every 1000th", 200_000, 1, 3, 8, 1, 1, 1, False,
                       mixed=(1000, wlsqm.b1_X), note=" (of 2M)")
    if want("cfg5"):
        n = int(1_000_000 * sc)
        s, xi, fi = run_expert("cfg5 cloud (cfg2)", n, 2, 4, 30, 0, 1, 1, False)
        nq = 16 * n
        I = torch.arange(n, device="cuda", dtype=torch.int64).repeat_interleave(16)
        g = torch.Generator(device="cuda").manual_seed(5)
        xq = xi[I] + 0.3e-2 * (2 * torch.rand((nq, 2), dtype=torch.float64, device="cuda", generator=g) - 1)
        s.tree = object()     # index given explicitly: no kd-tree needed
        ms = timeit(lambda: s.interpolate(xq, diff='all', I=I), reps=5)
        emit("cfg5 interpolate all 15 derivative slots, 16 queries per model (one pass, extension)", "interpolate", nq, ms,
             16 + 8 + 120 + 136 / 16)
        ms1 = timeit(lambda: s.interpolate(xq, diff=0, I=I), reps=5)
        emit("cfg5 interpolate diff=0 (one reference-style call)", "interpolate", nq, ms1, 16 + 8 + 8 + 136 / 16)

        def all15():
            for d in range(15):
                s.interpolate(xq, diff=d, I=I)
        ms15 = timeit(all15, reps=3, warm=1)
        emit("cfg5 interpolate d=0..14 (15 reference-style calls)", "interpolate", nq, ms15, 15 * (16 + 8 + 8 + 136 / 16))
        if cpu and reference() is not None:
            import os
            import time
            ref = reference()
            nc = 100_000
            gq = torch.Generator().manual_seed(5)
            xi_c, xk_c = (t.cpu() for t in make(nc, 2, 30))
            xi_h, xk_h = xi_c.numpy(), xk_c.numpy()
            fk_h = field(xk_c).numpy().copy()
            fi_h = np.zeros((nc, 15))
            nt = os.cpu_count() or 1
            sr = ref.ExpertSolver(2, np.full(nc, 30, np.int32), np.full(nc, 4, np.int32), np.zeros(nc, np.int64), np.full(nc, 1, np.int32),
                                  algorithm=1, do_sens=False, ntasks=nt)
            sr.prepare(xi_h, xk_h)
            sr.solve(fk_h, fi_h)
            Ih = np.repeat(np.arange(nc), 16)
            xq_h = xi_h[Ih] + 0.3e-2 * (2 * torch.rand((16 * nc, 2), dtype=torch.float64, generator=gq).numpy() - 1)
            t0 = time.perf_counter()
            sr.prep_interpolate()
            t_tree = time.perf_counter() - t0
            t0 = time.perf_counter()
            _, I0 = sr.interpolate(xq_h, mode='nearest', diff=0)          # first call: cKDTree.query inside
            t_first = time.perf_counter() - t0
            t0 = time.perf_counter()
            for d in range(15):
                sr.interpolate(xq_h, mode='nearest', diff=d, I=I0)
            t15 = time.perf_counter() - t0
            ex = {"ntasks": nt, "sample": "%d-point cloud, 16 queries per model (of 1M / 16M)" % nc}
            cpu_emit("cfg5 prep_interpolate (cKDTree build over the model origins)", "interpolate", nc, t_tree, ex)
            cpu_emit("cfg5 interpolate diff=0, first call (cKDTree.query for the model index inside)", "interpolate", 16 * nc, t_first, ex)
            cpu_emit("cfg5 interpolate d=0..14 (15 reference-style calls, I reused)", "interpolate", 16 * nc, t15, ex)


if __name__ == "__main__":
    main()
