#!/usr/bin/env python
"""Batched dense drivers: seconds per independent n x n problem, device-resident
batches, next to the reference's own drivers on this box's host cores and the figure the reference publishes
(README.md:85-99 / lapack_timings.png, read off a log-log plot: mgeneralp ~9e-7 s @ n=3, ~2e-6 @ n=15, ~8e-6 @ n=36).

    python benchmarks/lapack_bench.py [--nlhs 1000000]

Bytes per problem (HBM floor): dgesv reads A (8 n^2) + b (8 n), writes LU (8 n^2) + ipiv (4 n) + x (8 n).
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "python-wlsqm_b200"), str(ROOT / "oracle")]
from wlsqm_b200.utils import lapackdrivers as ld  # noqa: E402

PUBLISHED = {3: 9e-7, 15: 2e-6, 36: 8e-6}
PEAK = 6457.7
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def gpu_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nlhs", type=int, default=1_000_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scalers", action="store_true", help="also time the batched matrix scalers (mdo_rescale)")
    a = ap.parse_args()
    ref = None
    if not a.no_cpu:
        try:
            import oracle as orc
            if orc.load_reference() is not None:
                from wlsqm.utils import lapackdrivers as ref  # the unmodified reference (oracle/_ref)
        except Exception:
            ref = None
    cores = os.cpu_count() or 1
    g = torch.Generator(device="cuda").manual_seed(0)
    for n in (3, 15, 36):
        nlhs = a.nlhs if n < 36 else a.nlhs // 4
        # Fortran (n, n, nlhs) = transposed view of a C-contiguous (nlhs, n, n) tensor
        base = torch.randn((nlhs, n, n), dtype=torch.float64, device="cuda", generator=g)
        sym = 0.5 * (base + base.transpose(1, 2))
        rhs = torch.randn((nlhs, n), dtype=torch.float64, device="cuda", generator=g)
        for name, mat, solve in (("mgeneral (dgesv)", base, ld.mgeneral), ("msymmetric (dsysv)", sym, ld.msymmetric)):
            work_a, work_b = torch.empty_like(mat), torch.empty_like(rhs)

            def step():
                work_a.copy_(mat)
                work_b.copy_(rhs)
                solve(work_a.permute(2, 1, 0), work_b.t())

            def copies():
                work_a.copy_(mat)
                work_b.copy_(rhs)
            t = gpu_time(step) - gpu_time(copies)      # the driver works in place: the refresh copies are subtracted
            bytes_per = 16 * n * n + 20 * n
            line = {"driver": name, "n": n, "nlhs": nlhs, "s_per_problem": t / nlhs, "problems_per_s": nlhs / t,
                    "GBps": bytes_per * nlhs / t / 1e9, "hbm_frac": bytes_per * nlhs / t / 1e9 / PEAK,
                    "published_reference_s_per_problem": PUBLISHED.get(n) if "general" in name else None}
            if ref is not None:
                m = min(nlhs, 200_000 if n < 36 else 50_000)
                Ah = np.asfortranarray(mat[:m].cpu().numpy().transpose(2, 1, 0))
                bh = np.asfortranarray(rhs[:m].cpu().numpy().T)
                fn = ref.mgeneralp if "general" in name else ref.msymmetricp
                best = None
                for nt in sorted({1, cores}):
                    A2, b2 = Ah.copy(order="F"), bh.copy(order="F")
                    t0 = time.perf_counter()
                    fn(A2, b2, nt)
                    dt = time.perf_counter() - t0
                    if best is None or dt < best[0]:
                        best = (dt, nt)
                line["cpu_reference"] = {"s_per_problem": best[0] / m, "cores": best[1], "sample": m}
                line["speedup_vs_cpu_reference"] = (best[0] / m) / (t / nlhs)
            print(json.dumps(line), flush=True)
    if a.scalers:
        # batched equilibration (mdo_rescale) of Gram-like 15 x 15 matrices, every algorithm of ScalingAlgo; CPU: the
        # reference's do_rescale matrix by matrix (it has no batched scaler)
        n, nlhs = 15, a.nlhs // 4
        c = torch.randn((nlhs, 40, n), dtype=torch.float64, device="cuda", generator=g) * (10.0 ** torch.linspace(-3, 1, n, dtype=torch.float64, device="cuda"))
        gram = c.transpose(1, 2) @ c
        work = torch.empty_like(gram)
        for algo in ld.ScalingAlgo:
            def step():
                work.copy_(gram)
                ld.mdo_rescale(work.permute(2, 1, 0), algo)

            def copies():
                work.copy_(gram)
            t = gpu_time(step) - gpu_time(copies)
            bytes_per = 16 * n * n + 16 * n
            line = {"driver": "mdo_rescale " + algo.name, "n": n, "nlhs": nlhs, "s_per_problem": t / nlhs, "problems_per_s": nlhs / t,
                    "GBps": bytes_per * nlhs / t / 1e9, "hbm_frac": bytes_per * nlhs / t / 1e9 / PEAK}
            if ref is not None:
                m = 20_000
                Ah = gram[:m].cpu().numpy()
                t0 = time.perf_counter()
                for l in range(m):
                    ref.do_rescale(np.asfortranarray(Ah[l]), int(algo))
                dt = time.perf_counter() - t0
                line["cpu_reference"] = {"s_per_problem": dt / m, "cores": 1, "sample": m, "note": "do_rescale per matrix (no batched scaler in the reference)"}
                line["speedup_vs_cpu_reference"] = (dt / m) / (t / nlhs)
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
