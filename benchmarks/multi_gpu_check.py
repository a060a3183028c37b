#!/usr/bin/env python
"""Multi-GPU check under torchrun (one rank per GPU, NCCL): contiguous case ranges per rank, no data-path collective
in prepare / solve; with the hoods extension the per-point data f is all-gathered once per step.  Every rank compares
its rows with an unsharded solve of the whole batch done on its own GPU (bit for bit) and rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/multi_gpu_check.py
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "python-wlsqm_b200")]
import wlsqm_b200 as wlsqm                 # noqa: E402
from wlsqm_b200 import parallel            # noqa: E402
import workloads as wl                     # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, k, dim, order, no = 400_000, 30, 2, 4, 15
    x = wl.cloud(n, dim)                                      # same seed on every rank: the replicated cloud
    x_d = torch.from_numpy(x).to(dev)
    hoods_d = wlsqm.knn_hoods(x_d, k)
    f = torch.from_numpy(wl.field(x)).to(dev)
    meta = (np.full(n, k, np.int32), np.full(n, order, np.int32), np.zeros(n, np.int64), np.full(n, 1, np.int32))
    # unsharded result on this GPU
    full = wlsqm.ExpertSolver(dim, *meta, device=local)
    full.prepare_hoods(x_d, hoods_d)
    fi_full = torch.zeros((n, no), dtype=torch.float64, device=dev)
    full.solve_hoods(f, fi_full)
    # sharded: this rank's contiguous range, f given as the local slice (all-gathered inside)
    sh = parallel.ShardedExpertSolver(dim, *meta, rank=rank, world=world, device=local)
    sh.prepare_hoods(x_d, hoods_d)
    fi_loc = torch.zeros((sh.hi - sh.lo, no), dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    sh.solve_hoods(f[sh.lo:sh.hi].contiguous(), fi_loc, local=True, f_is_local=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(steps):
        sh.solve_hoods(f[sh.lo:sh.hi].contiguous(), fi_loc, local=True, f_is_local=True)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    same_rows = bool(torch.equal(fi_loc, fi_full[sh.lo:sh.hi]))
    whole = sh.gather(fi_loc)                                  # collective gather: the whole field on every rank
    same_all = bool(torch.equal(whole, fi_full))

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    f_full = f
    out_buf = torch.empty((n, no), dtype=torch.float64, device=dev)
    ms_local = timed(lambda: sh.solve_hoods(f_full, fi_loc, local=True))
    ms_nccl = timed(lambda: (sh.solve_hoods(f_full, fi_loc, local=True), sh.gather(fi_loc, out=out_buf)))
    # fused gather: the solve kernel stores its rows into every GPU's copy of the global array over NVLink
    fi_glob = sh.enable_fused_gather()
    sh.solve_hoods(f_full, fi_loc, local=True)
    sh.sync_gather()
    torch.cuda.synchronize()
    same_fused = bool(torch.equal(fi_glob, fi_full))
    ms_fused = timed(lambda: (sh.solve_hoods(f_full, fi_loc, local=True), sh.sync_gather()))
    sh.disable_fused_gather()
    ok = torch.tensor([int(same_rows and same_all and same_fused)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": "sharded solve_hoods == unsharded, bit for bit, on every rank (local rows, NCCL all-gather, "
                                   "fused peer-store gather)", "ok": bool(ok.item()),
                          "world": world, "points": n, "ms_per_step_incl_allgather_of_f": float(ms.item()),
                          "ms_per_step_local_rows_only": ms_local, "ms_per_step_with_nccl_all_gather_of_fi": ms_nccl,
                          "ms_per_step_with_fused_gather_of_fi": ms_fused,
                          "points_per_s_fused": n / (ms_fused * 1e-3)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok.item() else 1)


if __name__ == "__main__":
    main()
