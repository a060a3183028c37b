#!/usr/bin/env python
"""Host<->device ceiling of the box for the traffic pattern of one end-to-end solve step:
N processes (one per GPU), page-locked buffers, every step a concurrent H2D of 240 MB (fk of 1M points, k = 30) and
D2H of 120 MB (fi, 15 columns) per GPU on two streams -- NO kernel.  What bench.py's `e2e` can reach at N GPUs.

    python benchmarks/pcie_ceiling.py [--out profiles/r02_pcie_ceiling.json]     (on a box with >= 1 GPU; sweeps N = 1, 2, 4, 8)

Variants per N: plain page-locked buffers; the rank bound to the CPUs next to its GPU before allocating (NUMA-local
pages); write-combined input buffer on top of that.  The best variant per N is the ceiling (`per_n`).
"""
import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "python-wlsqm_b200")]

H2D_BYTES, D2H_BYTES = 240_000_000, 120_000_000


VARIANTS = ("pinned", "pinned+affinity", "wc+affinity")


def worker(rank, world, steps, barrier, q):
    import torch
    import wlsqm_b200 as wlsqm
    sys.path.insert(0, str(ROOT))
    import bench
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    d_in = torch.empty(H2D_BYTES // 8, dtype=torch.float64, device=dev)
    d_out = torch.ones(D2H_BYTES // 8, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for variant in VARIANTS:
        if variant != "pinned":
            bench.bind_to_gpu_cpus(rank)          # (stays bound for the remaining variants)
        src = wlsqm.pinned_empty((H2D_BYTES // 8,), write_combined=(variant == "wc+affinity"))
        src[...] = 1.0
        dst = wlsqm.pinned_empty((D2H_BYTES // 8,))
        dst[...] = 0.0
        src_t, dst_t = torch.from_numpy(src), torch.from_numpy(dst)
        for _ in range(3):
            with torch.cuda.stream(s1):
                d_in.copy_(src_t, non_blocking=True)
            with torch.cuda.stream(s2):
                dst_t.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        res = {}
        for what in ("both", "h2d", "d2h"):
            barrier.wait()
            t0 = time.perf_counter()
            for _ in range(steps):
                if what in ("both", "h2d"):
                    with torch.cuda.stream(s1):
                        d_in.copy_(src_t, non_blocking=True)
                if what in ("both", "d2h"):
                    with torch.cuda.stream(s2):
                        dst_t.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize()          # a step is complete when both copies are (like solve() on host arrays)
            dt = time.perf_counter() - t0
            barrier.wait()
            res[what] = 1e3 * dt / steps
        out[variant] = res
        del src_t, dst_t
        wlsqm.pinned_free(src)
        wlsqm.pinned_free(dst)
    q.put((rank, out))


def topo():
    out = {}
    for name, cmd in (("nvidia_smi_topo", ["nvidia-smi", "topo", "-m"]), ("lscpu", ["lscpu"]), ("numactl", ["numactl", "-H"])):
        try:
            out[name] = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout[-4000:]
        except Exception as exc:
            out[name] = repr(exc)
    return out


def main():
    import torch
    import torch.multiprocessing as mp
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    ngpu = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    result = {"what": "N ranks, each: concurrent H2D 240 MB + D2H 120 MB per step on two streams, page-locked buffers, no kernel",
              "gpus_on_box": ngpu, "host_cpus": os.cpu_count(), "variants": {}, "per_n": {}, "topology": topo()}
    for n in (1, 2, 4, 8):
        if n > ngpu:
            break
        barrier, q = ctx.Barrier(n), ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, n, a.steps, barrier, q)) for r in range(n)]
        for p in procs:
            p.start()
        res = [q.get(timeout=900) for _ in procs]
        for p in procs:
            p.join(timeout=120)
        for variant in VARIANTS:
            ms = {w: max(r[1][variant][w] for r in res) for w in ("both", "h2d", "d2h")}
            entry = {"ms_per_step": ms["both"], "h2d_only_ms": ms["h2d"], "d2h_only_ms": ms["d2h"],
                     "h2d_GBps_aggregate_concurrent": n * H2D_BYTES / (ms["both"] * 1e-3) / 1e9,
                     "d2h_GBps_aggregate_concurrent": n * D2H_BYTES / (ms["both"] * 1e-3) / 1e9,
                     "h2d_GBps_per_gpu_alone_direction": H2D_BYTES / (ms["h2d"] * 1e-3) / 1e9,
                     "d2h_GBps_per_gpu_alone_direction": D2H_BYTES / (ms["d2h"] * 1e-3) / 1e9,
                     "points_per_s_ceiling": n * 1_000_000 / (ms["both"] * 1e-3)}
            result["variants"].setdefault(str(n), {})[variant] = entry
            print(json.dumps({"n": n, "variant": variant, **entry}), flush=True)
        best = min(result["variants"][str(n)].items(), key=lambda kv: kv[1]["ms_per_step"])
        result["per_n"][str(n)] = dict(best[1], variant=best[0])
    if a.out:
        Path(a.out).write_text(json.dumps(result, indent=1))
    print(json.dumps({"per_n": result["per_n"]}), flush=True)


if __name__ == "__main__":
    main()
