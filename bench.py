#!/usr/bin/env python
"""bench.py -- the headline benchmark: ExpertSolver.solve on the 1M-point 2D order-4 (15 DOF, k=30) stream.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n POINTS_PER_GPU]

BASELINE.json metric: "local fits/s and ExpertSolver.solve pts/s (2D order-4, 1M pts) vs FP64/HBM roofline";
configs[1]: "ExpertSolver 2D order-4 (15 DOF), 1M points, k=30, prepare once then solve() time steps with
varying fk".  One *step* = one solve() over the whole cloud with a fresh fk.

  value      points/s, whole job, fk/fi resident in HBM (CUDA tensors handed to the public API, zero copy);
             CUDA events on the stream the kernel runs on; max over ranks.
  e2e        the same call with HOST (pinned) numpy arrays: H2D of fk and D2H of fi inside the timed region.
  roofline   dominant kernel = solve_kernel: algorithmic bytes/point = 8*(nr*nk + nr*n_known + nk + 2*no)
             (4080 B for this config) / mean launch duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the unmodified reference (oracle/_ref, OpenMP) on a bounded sample of the same workload,
             on this box's host cores (rank 0, N=1 only).
  prepare    fits/s of ExpertSolver.prepare on the same cloud and its FP64 flop rate (extra keys).
  gather     the step with the GLOBAL fi assembled on every rank: fused (peer stores from the solve kernel over
             NVLink + one 4-byte all-reduce) and with an NCCL all-gather after the kernel.
  strong     fixed total problems split N ways (extra keys): cfg2 at 1M points, cfg3 (3D order 4, 4M points, k=60,
             ALGO_ITERATIVE(3), do_sens) with fi gathered inside the timed region and sens left sharded.

Multi-GPU (torchrun, one rank per GPU): ONE global cloud of N x --n points (same seed on every rank); rank r owns the
contiguous case range [r n, (r+1) n) through wlsqm_b200.parallel.ShardedExpertSolver.  prepare / solve need no
data-path collective; weak scaling (every rank owns --n points); value = all points / max-over-ranks time.

--impl reference times the reference's own CPU implementation of the same path on the host cores: the same --n-point
workload as ExpertSolver chunks of <= 250k cases (its arena is sized in C int bytes: one solver holds at most ~343k
cases of this configuration), step time = sum over the chunks; best of the two reference builds
(generic -O2 / tuned) and of ntasks in {1, 8, 16, cores}.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for _p in (ROOT, ROOT / "python-wlsqm_b200"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import workloads as wl  # noqa: E402

DIM, ORDER, K, KNOWNS = 2, 4, 30, 0
NO = 15
NR = NO - bin(KNOWNS).count("1")
BYTES_PER_POINT = 8 * (NR * K + NR * (NO - NR) + K + 2 * NO)          # 4080
FLOPS_PREP = 33540                                                     # SURVEY.md 8d, cfg2 knowns=0 incl. operator
METRIC = "ExpertSolver.solve points/s (2D order-4, 15 DOF, k=30, 1M-point cloud)"
UNIT = "points/s"
REF_CHUNK = 250_000                                                    # cfg2 <= 300k cases per reference solver
# cfg3 (BASELINE.json configs[2]): 3D order 4, k = 60, b3_F, ALGO_ITERATIVE(3), do_sens
C3_BYTES = 8 * (34 * 60 + 34 * 1 + 60 + 2 * 35) + 8 * (60 * 3 + 3) + 8 * 60 * 35      # 35 896


def config_dict(n):
    """what both arms (--impl b200 / reference) run: identical by construction"""
    return {"workload": "cfg2: ExpertSolver 2D order-4 k=30 knowns=0 WEIGHT_UNIFORM ALGO_BASIC, "
                        "prepare once then solve() per step with varying fk",
            "points_per_gpu": n, "cloud": "fixed-density random cloud (h0 = 1e-2), seed 42, k-nearest-neighbour hoods",
            "sharding": "N GPUs: one global cloud of N x points_per_gpu points, contiguous case ranges per GPU, "
                        "no data-path collective in solve",
            "l2": "inputs larger than L2: operators 3.6 GB + rotating fk buffers of 240 MB each per step",
            "bytes_per_point": BYTES_PER_POINT}


def _fp64_peak():
    """Measured FP64 vector-FMA peak of this pool's B200 (benchmarks/fp64_peak.cu; MEASURED_PEAKS.json has no FP64 figure)."""
    p = ROOT / "profiles" / "fp64_peak.json"
    try:
        j = json.loads(p.read_text())
        return float(j["dfma_tflops"]), float(j["dmma_m8n8k4_tflops"]), "measured (profiles/fp64_peak.json, benchmarks/fp64_peak.cu)"
    except Exception:
        return 37.0, 37.0, "nominal (B200 datasheet FP64)"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _pcie_ceiling(world):
    """measured host<->device ceiling of the 8-GPU box for this many concurrent ranks (benchmarks/pcie_ceiling.py)"""
    try:
        j = json.loads((ROOT / "profiles" / "r02_pcie_ceiling.json").read_text())
        return j["per_n"].get(str(world))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def count(self):
        return len(self.lines)

    def wait_ready(self, timeout=3.0):
        """block until nvidia-smi has produced its first line (its start-up takes longer than the timed region)"""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, first=0):
        """summary of the samples from line `first` on (the ones taken under the load of the timed steps)"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[first:]:
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(n: int, seed: int):
    x = wl.cloud(n, DIM, seed=seed)
    hoods = wl.hoods_knn(x, K)
    f = wl.field(x)
    return x, hoods, f


def meta(n):
    return (np.full(n, K, np.int32), np.full(n, ORDER, np.int32), np.full(n, KNOWNS, np.int64),
            np.full(n, 1, np.int32))   # WEIGHT_UNIFORM


def bind_to_gpu_cpus(gpu_index):
    """run this rank on the CPUs next to its GPU (NUMA-local page-locked buffers for the host<->device legs); returns the
    previous affinity so that the CPU-baseline leg can have all cores back"""
    prev = None
    try:
        prev = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (max(prev) // 64) + 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & prev
        if cpus and len(cpus) < len(prev):
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    return prev


# ------------------------------------------------------------------------------------------------------
# the reference arm
# ------------------------------------------------------------------------------------------------------
def _ref_chunks(n):
    nch = max(1, -(-n // REF_CHUNK))
    return [(c * n // nch, (c + 1) * n // nch) for c in range(nch)]


def _ref_probe(ref, x, xk, fks, lo, hi, candidates, steps=3):
    """points/s of solve() on one chunk for every ntasks candidate"""
    out = {}
    nk, od, kn, wm = meta(hi - lo)
    fi = np.zeros((hi - lo, NO))
    for nt in candidates:
        s = ref.ExpertSolver(DIM, nk, od, kn, wm, algorithm=ref.ALGO_BASIC, do_sens=False, ntasks=nt)
        s.prepare(x[lo:hi], xk[lo:hi])
        s.solve(fks[0][lo:hi], fi)
        t0 = time.perf_counter()
        for t in range(steps):
            s.solve(fks[t % len(fks)][lo:hi], fi)
        out[nt] = (hi - lo) * steps / (time.perf_counter() - t0)
        del s
    return out


def run_reference(args, rank, world):
    """The reference's own CPU path (oracle/_ref): the same --n-point workload as chunked ExpertSolvers."""
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as orc
    cores = os.cpu_count() or 1
    variants = orc.reference_variants()
    n = args.n
    chunks = _ref_chunks(n)
    cands = sorted({1, 8, 16, cores} & set(range(1, cores + 1))) or [1]
    if args.probe_variant:          # child process: probe one build, print {ntasks: points/s}
        ref = orc.load_reference(args.probe_variant)
        lo, hi = chunks[0]
        x, hoods, f = build_workload(hi - lo, 42)
        xk = np.ascontiguousarray(x[hoods])
        fks = [np.ascontiguousarray(wl.field_step(f, t)[hoods]) for t in range(2)]
        print(json.dumps({str(k): v for k, v in _ref_probe(ref, x, xk, fks, 0, hi - lo, cands).items()}), flush=True)
        return
    probes = {}
    for v in variants:
        try:
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--probe-variant", v,
                                "--n", str(n)], capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
            probes[v] = {int(k): float(p) for k, p in json.loads(r.stdout.strip().splitlines()[-1]).items()}
        except Exception as exc:
            print("probe of the %s reference build failed: %r" % (v, exc), file=sys.stderr)
    best_v, best_nt = "generic", cands[-1]
    if probes:
        best_v, best_nt = max(((v, nt) for v in probes for nt in probes[v]), key=lambda t: probes[t[0]][t[1]])
    ref = orc.load_reference(best_v) if variants else None
    x, hoods, f = build_workload(n, 42)
    xk = np.ascontiguousarray(x[hoods])
    fks = [np.ascontiguousarray(wl.field_step(f, t)[hoods]) for t in range(2)]
    fi = np.zeros((n, NO))
    if ref is not None:
        kind = "reference"
        solvers = []
        for lo, hi in chunks:
            nk, od, kn, wm = meta(hi - lo)
            s = ref.ExpertSolver(DIM, nk, od, kn, wm, algorithm=ref.ALGO_BASIC, do_sens=False, ntasks=best_nt)
            s.prepare(x[lo:hi], xk[lo:hi])
            solvers.append(s)

        def step(t):
            for s, (lo, hi) in zip(solvers, chunks):
                s.solve(fks[t % 2][lo:hi], fi[lo:hi])
    else:
        kind, best_nt = "port", 1
        nk, od, kn, wm = meta(n)
        so = orc.OracleSolver(DIM, nk, od, kn, wm)
        so.prepare(x, xk)

        def step(t):
            so.solve(fks[t % 2], fi)
    for w in range(args.warmup):
        step(w)
    t0 = time.perf_counter()
    for t in range(args.steps):
        step(t)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(n),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": best_nt, "kind": kind, "host_cores": cores,
                         "sample": f"{n} points per step as {len(chunks)} ExpertSolver chunks of <= {REF_CHUNK} cases "
                                   f"(step = sum over the chunks), {args.steps} steps; build '{best_v}', ntasks={best_nt}: "
                                   f"best of builds {sorted(probes)} x ntasks {cands}",
                         "probe_points_per_s": {v: {str(k): p for k, p in d.items()} for v, d in probes.items()}},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    """bounded sample of the same workload on the host cores: ~10-30 s of CPU work"""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as orc
    ref = orc.load_reference()
    cores = os.cpu_count() or 1
    n = min(args.n, 100_000)
    x, hoods, f = build_workload(n, 42)
    nk, od, kn, wm = meta(n)
    xk = np.ascontiguousarray(x[hoods])
    fks = [np.ascontiguousarray(wl.field_step(f, t)[hoods]) for t in range(2)]
    fi = np.zeros((n, NO))
    steps = 10
    out = {}
    if ref is not None:
        best = None
        for nt in sorted({1, 8, 16, cores} & set(range(1, cores + 1))):
            s = ref.ExpertSolver(DIM, nk, od, kn, wm, algorithm=ref.ALGO_BASIC, do_sens=False, ntasks=nt)
            t0 = time.perf_counter()
            s.prepare(x, xk)
            tp = time.perf_counter() - t0
            s.solve(fks[0], fi)
            t0 = time.perf_counter()
            for t in range(steps):
                s.solve(fks[t % 2], fi)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, nt, tp)
            del s
        dt, nt, tp = best
        out = {"value": n * steps / dt, "unit": UNIT, "cores": nt, "kind": "reference", "host_cores": cores,
               "sample": f"{n} points x {steps} solve() steps (prepare {n / tp:.3g} fits/s at ntasks={nt}); "
                         f"best of ntasks in {{1, 8, 16, {cores}}}",
               "prepare_fits_per_s": n / tp}
        try:    # configs[0]: the reference's one-shot fit_2D_many_parallel, 10k points, order 2, k=12
            n1, k1 = 10_000, 12
            x1 = wl.cloud(n1, 2, unit_box=True)
            h1 = wl.hoods_knn(x1, k1)
            f1 = wl.field(x1)
            xk1, fk1 = np.ascontiguousarray(x1[h1]), np.ascontiguousarray(f1[h1])
            fi1 = np.zeros((n1, 6))
            fi1[:, 0] = f1
            m1 = (np.full(n1, k1, np.int32), np.full(n1, 2, np.int32), np.full(n1, ref.b2_F, np.int64),
                  np.full(n1, ref.WEIGHT_CENTER, np.int32))
            tb = None
            for nt1 in sorted({1, 8, cores}):
                for _ in range(3):
                    t0 = time.perf_counter()
                    ref.fit_2D_many_parallel(xk1, fk1, m1[0], x1, fi1, None, 0, m1[1], m1[2], m1[3], ntasks=nt1)
                    d = time.perf_counter() - t0
                    if tb is None or d < tb[0]:
                        tb = (d, nt1)
            out["cfg1_fit_many_fits_per_s"] = n1 / tb[0]
            out["cfg1_fit_many_ntasks"] = tb[1]
        except Exception as exc:
            print("cpu cfg1 leg failed: %r" % (exc,), file=sys.stderr)
    else:
        s = orc.OracleSolver(DIM, nk, od, kn, wm)
        t0 = time.perf_counter()
        s.prepare(x, xk)
        tp = time.perf_counter() - t0
        t0 = time.perf_counter()
        for t in range(steps):
            s.solve(fks[t % 2], fi)
        dt = time.perf_counter() - t0
        out = {"value": n * steps / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{n} points x {steps} solve() steps, scalar C port", "prepare_fits_per_s": n / tp}
    return out


# ------------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------------
def strong_leg(torch, dist, wlsqm, parallel, rank, world, local_rank, name, n_total, dim, order, k, knowns, wmeth, algo,
               do_sens, bytes_per_point, steps, warm):
    """a FIXED total problem split over the ranks (contiguous ranges of one cloud): per step every rank solves its range
    and the global fi is assembled on every rank; sens stays sharded.  Times (max over ranks, CUDA events):
    local rows only / + fused gather (peer stores + 4-byte all-reduce) / + NCCL all-gather after the kernel."""
    dev = torch.device("cuda", local_rank)
    no = wlsqm.number_of_dofs(dim, order)
    x = wl.cloud(n_total, dim)
    x_d = torch.from_numpy(x).to(dev)
    hoods_d = wlsqm.knn_hoods(x_d, k)
    f_d = torch.from_numpy(wl.field(x)).to(dev)
    m = (np.full(n_total, k, np.int32), np.full(n_total, order, np.int32), np.full(n_total, knowns, np.int64),
         np.full(n_total, wmeth, np.int32))
    sh = parallel.ShardedExpertSolver(dim, *m, rank=rank, world=world, device=local_rank, algorithm=algo, do_sens=do_sens,
                                      max_iter=3)
    lo, hi = sh.lo, sh.hi
    nloc = hi - lo
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sh.prepare_hoods(x_d, hoods_d)
    torch.cuda.synchronize()
    e0.record()
    sh.prepare_hoods(x_d, hoods_d)
    e1.record()
    torch.cuda.synchronize()
    prep_ms = e0.elapsed_time(e1)
    hl = hoods_d[lo:hi].long()
    fk = [f_d[hl].contiguous(), (1.5 * f_d)[hl].contiguous()]
    del hl, hoods_d
    fi_loc = torch.zeros((nloc, no), dtype=torch.float64, device=dev)
    fi_loc[:, 0] = f_d[lo:hi]
    sens = torch.empty((nloc, k, no), dtype=torch.float64, device=dev) if do_sens else None
    # the geometry was gathered by prepare_hoods; the per-step data comes pre-gathered (the reference's signature)
    s = sh.solver

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn):
        for w in range(warm):
            fn(w)
        barrier()
        e0.record()
        for t in range(steps):
            fn(t)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms
    ms_local = timed(lambda t: s.solve(fk[t % 2], fi_loc, sens))
    out_buf = torch.empty((n_total, no), dtype=torch.float64, device=dev)
    ms_nccl = timed(lambda t: (s.solve(fk[t % 2], fi_loc, sens), sh.gather(fi_loc, out=out_buf))) if dist is not None else ms_local
    fused_error, same, ms_fused = None, None, ms_nccl
    try:
        fi_glob = sh.enable_fused_gather()
    except Exception as exc:          # (e.g. CUDA IPC not permitted in this container: the collective gather is the step then)
        fused_error = repr(exc)[:200]
    if fused_error is None:
        ms_fused = timed(lambda t: (s.solve(fk[t % 2], fi_loc, sens), sh.sync_gather()))
        torch.cuda.synchronize()
        same = bool(torch.equal(fi_glob[lo:hi], fi_loc))
        if dist is not None:
            s.solve(fk[0], fi_loc, sens)
            sh.gather(fi_loc, out=out_buf)
            s.solve(fk[0], fi_loc, sens)
            sh.sync_gather()
            torch.cuda.synchronize()
            same = same and bool(torch.equal(fi_glob, out_buf))
            ok = torch.tensor([int(same)], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            same = bool(ok.item())
    sh.close()
    del sens, fk, out_buf, fi_loc
    torch.cuda.empty_cache()
    wlsqm.pool_trim()
    peak = _peaks()[0]
    return {"workload": name, "points_total": n_total, "points_per_gpu": nloc, "n_gpus": world,
            "ms_per_step": ms_fused, "points_per_s": n_total / (ms_fused * 1e-3),
            "ms_per_step_local_rows_only": ms_local, "ms_per_step_nccl_all_gather_after_kernel": ms_nccl,
            "gather": "fused: the solve kernel stores every row into all ranks' copies of the global fi (peer memory over "
                      "NVLink) + one 4-byte NCCL all-reduce per step; sens stays sharded",
            "fused_equals_nccl_gather_bit_for_bit": same, "fused_gather_error": fused_error,
            "prepare_ms": prep_ms, "bytes_per_point": bytes_per_point,
            "hbm_frac_per_gpu": bytes_per_point * nloc / (ms_fused * 1e-3) / 1e9 / peak}


def run_b200(args, rank, world, local_rank):
    import torch
    import wlsqm_b200 as wlsqm
    from wlsqm_b200 import parallel

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: wlsqm_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = bind_to_gpu_cpus(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    n_total = world * n
    # ONE global cloud (same seed on every rank); neighbourhoods by the device-side search (== cKDTree, tests/test_gpu_grid.py)
    x = wl.cloud(n_total, DIM, seed=42)
    f = wl.field(x)
    x_d = torch.from_numpy(x).to(dev)
    hoods_all = wlsqm.knn_hoods(x_d, K)
    sh = parallel.ShardedExpertSolver(DIM, *meta(n_total), rank=rank, world=world, device=local_rank,
                                      algorithm=wlsqm.ALGO_BASIC, do_sens=False, ntasks=1)
    lo, hi = sh.lo, sh.hi
    assert hi - lo == n
    hoods_loc32 = hoods_all[lo:hi].contiguous()
    hoods_d = hoods_loc32.long()
    del hoods_all
    xi_d = x_d[lo:hi]
    xk_d = x_d[hoods_d]                                   # (n, K, 2), the caller-side gather of the examples
    NBUF = 4                                              # distinct fk buffers, each 240 MB > L2 (126 MB)
    fk_d = []
    for t in range(NBUF):
        ft = torch.from_numpy(wl.field_step(f, t)).to(dev)
        fk_d.append(ft[hoods_d].contiguous())
    fi_d = torch.zeros((n, NO), dtype=torch.float64, device=dev)
    del hoods_d
    s = sh.solver
    nk, od, kn, wm = meta(n)

    # ---- prepare (timed separately; reported as fits/s) ----
    sh.prepare(xi_d, xk_d, local=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prep_ms = []
    for _ in range(3):
        e0.record()
        sh.prepare(xi_d, xk_d, local=True)
        e1.record()
        torch.cuda.synchronize()
        prep_ms.append(e0.elapsed_time(e1))
    prep_ms = min(prep_ms)

    # ---- one-shot fits (fit_2D_many_parallel, "local fits/s"): configs[0] shape and the headline cloud ----
    oneshot = None
    try:
        def wall(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            return float(np.median(ts))
        n1, k1 = 10_000, 12
        x1 = wl.cloud(n1, 2, unit_box=True)
        h1 = wl.hoods_knn(x1, k1)
        f1 = wl.field(x1)
        xk1, fk1 = np.ascontiguousarray(x1[h1]), np.ascontiguousarray(f1[h1])
        fi1 = np.zeros((n1, 6))
        fi1[:, 0] = f1
        m1 = (np.full(n1, k1, np.int32), np.full(n1, 2, np.int32), np.full(n1, wlsqm.b2_F, np.int64),
              np.full(n1, wlsqm.WEIGHT_CENTER, np.int32))
        t_host = wall(lambda: wlsqm.fit_2D_many_parallel(xk1, fk1, m1[0], x1, fi1, None, 0, m1[1], m1[2], m1[3], ntasks=8))
        d1 = [torch.from_numpy(a).to(dev) for a in (xk1, fk1, x1, fi1)]
        t_dev = wall(lambda: wlsqm.fit_2D_many_parallel(d1[0], d1[1], m1[0], d1[2], d1[3], None, 0, m1[1], m1[2], m1[3], ntasks=8))
        fi_big = torch.zeros((n, NO), dtype=torch.float64, device=dev)
        t_big = wall(lambda: wlsqm.fit_2D_many_parallel(xk_d, fk_d[0], nk, xi_d, fi_big, None, 0, od, kn, wm, ntasks=8), reps=3)
        oneshot = {"cfg1": {"workload": "configs[0]: fit_2D_many_parallel, 10k points, order 2, k=12, b2_F, WEIGHT_CENTER",
                            "host_arrays_ms": 1e3 * t_host, "host_arrays_fits_per_s": n1 / t_host,
                            "cuda_tensors_ms": 1e3 * t_dev, "cuda_tensors_fits_per_s": n1 / t_dev},
                   "headline_cloud": {"workload": "fit_2D_many_parallel on this rank's 1M points (order 4, k=30), CUDA tensors, "
                                                  "wall clock per call", "ms": 1e3 * t_big, "fits_per_s": n / t_big}}
        del fi_big, d1
    except Exception as exc:      # extra information must never break the headline line
        print("one-shot leg failed: %r" % (exc,), file=sys.stderr)
    del xk_d

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: device-resident solve steps ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    for w in range(args.warmup):
        s.solve(fk_d[w % NBUF], fi_d)
    barrier()
    first_sample = sampler.count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for t in range(args.steps):
        s.solve(fk_d[t % NBUF], fi_d)
        evs[t + 1].record()
    barrier()
    clocks = None
    if rank == 0:
        # the timed region (steps x 0.6 ms) is shorter than nvidia-smi's sampling period (25 ms): the same steps keep
        # running, untimed, until five samples have been taken under that load
        in_region = sampler.count() - first_sample
        t_end = time.time() + 2.0
        while sampler.count() - first_sample < 5 and time.time() < t_end:
            for t in range(20):
                s.solve(fk_d[t % NBUF], fi_d)
            torch.cuda.synchronize()
        clocks = sampler.stop(first_sample)
        clocks["window"] = ("%d sample(s) inside the timed region of %d steps, the rest while the same steps kept running untimed "
                            "right after it (sampling period 25 ms)" % (in_region, args.steps))
    barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    kern_ms = float(np.mean(per_launch_ms))

    # ---- the same steps with the GLOBAL fi assembled on every rank ----
    gather = None
    try:
        out_buf = torch.empty((n_total, NO), dtype=torch.float64, device=dev)
        gsteps = max(3, min(args.steps, 10))

        def run(fn):
            for w in range(3):
                fn(w)
            barrier()
            e0.record()
            for t in range(gsteps):
                fn(t)
            e1.record()
            barrier()
            return e0.elapsed_time(e1) / gsteps
        ms_nccl = run(lambda t: (s.solve(fk_d[t % NBUF], fi_d), sh.gather(fi_d, out=out_buf))) if dist is not None else None
        fi_glob = sh.enable_fused_gather()
        ms_fused = run(lambda t: (s.solve(fk_d[t % NBUF], fi_d), sh.sync_gather()))
        torch.cuda.synchronize()
        same = bool(torch.equal(fi_glob[lo:hi], fi_d))
        if dist is not None:
            s.solve(fk_d[0], fi_d)
            sh.gather(fi_d, out=out_buf)
            sh.sync_gather()
            torch.cuda.synchronize()
            same = same and bool(torch.equal(fi_glob, out_buf))
        sh.disable_fused_gather()
        gather = [ms_fused, ms_nccl or 0.0, 1.0 if same else 0.0]
        del out_buf
    except Exception as exc:
        print("gather leg failed: %r" % (exc,), file=sys.stderr)
    # ---- extension: the same steps without the solver's own copy of the solution (keep_solution(False): the reference's
    #      Case_set_fi copy, needed only by interpolate) ----
    nocopy_ms = None
    try:
        s.keep_solution(False)
        for w in range(3):
            s.solve(fk_d[w % NBUF], fi_d)
        barrier()
        e0.record()
        for t in range(args.steps):
            s.solve(fk_d[t % NBUF], fi_d)
        e1.record()
        barrier()
        nocopy_ms = e0.elapsed_time(e1) / args.steps
    except Exception as exc:
        print("keep_solution leg failed: %r" % (exc,), file=sys.stderr)
    finally:
        s.keep_solution(True)
    # the last device-resident state of fi for the e2e check below
    s.solve(fk_d[(args.steps - 1) % NBUF], fi_d)
    torch.cuda.synchronize()

    # ---- e2e: same public call, host (pinned) numpy arrays, H2D + D2H inside the timed region ----
    fk_h = [wlsqm.pinned_empty((n, K)) for _ in range(2)]
    for t in range(2):
        fk_h[t][...] = fk_d[t].cpu().numpy()
    fi_h = wlsqm.pinned_empty((n, NO))
    fi_h[...] = 0.0
    e2e_steps = max(3, min(args.steps, 10))
    for w in range(2):
        s.solve(fk_h[w % 2], fi_h)
    barrier()
    e0.record()
    for t in range(e2e_steps):
        s.solve(fk_h[t % 2], fi_h)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    # the same call with ORDINARY (pageable) numpy arrays, what a reference user passes without thinking about it
    page_ms = None
    try:
        if world > 1:
            raise RuntimeError("skip")          # single-GPU figure (host threads of N ranks would share the cores)
        fk_pg = [np.array(fk_h[t]) for t in range(2)]
        fi_pg = np.zeros((n, NO))
        s.solve(fk_pg[0], fi_pg)
        barrier()
        t0 = time.perf_counter()
        for t in range(3):
            s.solve(fk_pg[t % 2], fi_pg)
        page_ms = 1e3 * (time.perf_counter() - t0) / 3
        del fk_pg, fi_pg
    except Exception as exc:
        if world == 1:
            print("pageable leg failed: %r" % (exc,), file=sys.stderr)
    del fk_d

    # ---- extension: the same step fed per point (solve_hoods): this rank's slice of f in, fi out; the ranks all-gather f
    #      over NCCL (8 B per point) and the gather f[hoods] runs on the GPU ----
    hoods_ms = None
    try:
        s2 = wlsqm.ExpertSolver(DIM, nk, od, kn, wm, algorithm=wlsqm.ALGO_BASIC, do_sens=False, ntasks=1, device=local_rank)
        s2.prepare_hoods(x_d, hoods_loc32, xi=xi_d)
        f_h = [wlsqm.pinned_empty((n,)) for _ in range(2)]
        for t in range(2):
            f_h[t][...] = wl.field_step(f, t)[lo:hi]
        f_t = [torch.from_numpy(a) for a in f_h]
        f_loc_d = torch.empty((n,), dtype=torch.float64, device=dev)
        f_all_d = torch.empty((n_total,), dtype=torch.float64, device=dev)

        def hstep(t):
            f_loc_d.copy_(f_t[t % 2], non_blocking=True)
            if dist is not None:
                dist.all_gather_into_tensor(f_all_d, f_loc_d)
                s2.solve_hoods(f_all_d, fi_h)
            else:
                s2.solve_hoods(f_loc_d, fi_h)
        for w in range(2):
            hstep(w)
        barrier()
        e0.record()
        for t in range(e2e_steps):
            hstep(t)
        e1.record()
        barrier()
        hoods_ms = e0.elapsed_time(e1)
        del s2
    except Exception as exc:      # the extension must never break the headline line
        print("solve_hoods leg failed: %r" % (exc,), file=sys.stderr)

    # ---- max over ranks ----
    if dist is not None:
        g = gather or [0.0, 0.0, 0.0]
        tt = torch.tensor([total_ms, e2e_ms, prep_ms, kern_ms, hoods_ms or 0.0, g[0], g[1], -g[2], nocopy_ms or 0.0],
                          dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, prep_ms, kern_ms, hm, g0, g1, g2, ncm = (float(v) for v in tt.tolist())
        hoods_ms = hm if hoods_ms else None
        nocopy_ms = ncm if nocopy_ms else None
        if gather:
            gather = [g0, g1, -g2]
    sh.close()
    del x_d, xi_d, fi_d, hoods_loc32
    torch.cuda.empty_cache()
    wlsqm.pool_trim()

    # ---- strong scaling: fixed total problems split over the ranks (extra keys) ----
    strong = {}
    if not args.no_strong:
        for name, leg in (
                ("cfg2_1M", dict(name="cfg2: 2D order 4, k=30, knowns=0, WEIGHT_UNIFORM, ALGO_BASIC, 1M points TOTAL",
                                 n_total=1_000_000, dim=2, order=4, k=30, knowns=0, wmeth=1, algo=1, do_sens=False,
                                 bytes_per_point=BYTES_PER_POINT, steps=10, warm=3)),
                ("cfg3_4M", dict(name="cfg3: 3D order 4, k=60, b3_F, WEIGHT_CENTER, ALGO_ITERATIVE(3), do_sens, 4M points TOTAL",
                                 n_total=4_000_000, dim=3, order=4, k=60, knowns=1, wmeth=2, algo=2, do_sens=True,
                                 bytes_per_point=C3_BYTES, steps=5, warm=2))):
            try:
                if name == "cfg3_4M":
                    free, _ = torch.cuda.mem_get_info()
                    need = 165e9 / world + 8e9
                    if free < need:
                        strong[name] = {"skipped": "needs %.0f GB free per GPU, have %.0f" % (need / 1e9, free / 1e9)}
                        continue
                strong[name] = strong_leg(torch, dist, wlsqm, parallel, rank, world, local_rank, **leg)
            except Exception as exc:
                strong[name] = {"failed": repr(exc)[:300]}
                print("strong leg %s failed: %r" % (name, exc), file=sys.stderr)
                torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_src = _peaks()
        achieved = BYTES_PER_POINT * n / (kern_ms * 1e-3) / 1e9
        e2e_val = world * n * e2e_steps / (e2e_ms * 1e-3)
        e2e = {"value": e2e_val, "unit": UNIT,
               "h2d_bytes_per_step": int(n * K * 8), "d2h_bytes_per_step": int(n * NO * 8),
               "steps": e2e_steps, "host_buffers": "pinned, allocated on the CPUs next to the rank's GPU",
               "pcie_h2d_GBps_per_gpu": n * K * 8 / (e2e_ms * 1e-3 / e2e_steps) / 1e9,
               "note": "bound by the host->device copy of fk (240 MB per step and GPU), which the reference API makes part of every step"}
        ceil = _pcie_ceiling(world)
        if ceil:
            # the box's measured ceiling for this step's traffic pattern: N ranks, 240 MB in + 120 MB out each, no kernel
            ceil_pts = world * n / (ceil["ms_per_step"] * 1e-3)
            e2e["ceiling"] = {"points_per_s": ceil_pts, "ms_per_step": ceil["ms_per_step"], "frac": e2e_val / ceil_pts,
                              "source": "profiles/r02_pcie_ceiling.json (benchmarks/pcie_ceiling.py: copies only, no kernel)"}
        line = {
            "metric": METRIC, "value": world * n * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(n),
            "e2e": e2e,
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "wlsqm::solve_kernel<1,false,false,true>", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_POINT * n, "launch_ms": kern_ms},
            "prepare": {"fits_per_s": world * n / (prep_ms * 1e-3), "ms": prep_ms, "flops_per_fit": FLOPS_PREP,
                        "roofline": {"bound": "fp64", "achieved": FLOPS_PREP * n / (prep_ms * 1e-3) / 1e12,
                                     "peak": _fp64_peak()[0], "unit": "TFLOP/s",
                                     "frac": FLOPS_PREP * n / (prep_ms * 1e-3) / 1e12 / _fp64_peak()[0],
                                     "peak_dmma": _fp64_peak()[1], "peak_source": _fp64_peak()[2],
                                     "kernel": "wlsqm::prepare_reg_kernel<2,4>"}},
            "clocks": clocks,
        }
        if gather:
            line["gather"] = {"what": "the headline step with the GLOBAL fi (%d x 15) assembled on every rank" % n_total,
                              "fused_ms_per_step": gather[0], "fused_points_per_s": world * n / (gather[0] * 1e-3),
                              "nccl_all_gather_ms_per_step": gather[1] or None,
                              "fused_equals_nccl_bit_for_bit": bool(gather[2] > 0.5),
                              "how": "fused: solve_kernel stores every row into all ranks' copies over NVLink peer memory + a "
                                     "4-byte all-reduce per step; nccl: all_gather_into_tensor after the kernel"}
        if strong:
            line["strong"] = strong
        if page_ms:
            line["e2e_pageable_numpy"] = {"value": n / (page_ms * 1e-3), "unit": UNIT, "ms_per_step": page_ms,
                                          "note": "rank 0, ordinary numpy arrays: host threads copy through page-locked rings "
                                                  "(csrc/wlsqm_host.cu); cudaMemcpy from pageable memory gave 29 ms per step"}
        if oneshot:
            line["one_shot_fits"] = oneshot
        if nocopy_ms:
            line["solve_without_solution_copy"] = {
                "ms_per_step": nocopy_ms, "value": world * n / (nocopy_ms * 1e-3), "unit": UNIT,
                "note": "ExpertSolver.keep_solution(False): the solver's own copy of fi (the reference's Case_set_fi, read only by "
                        "interpolate) is not written; 8*no bytes per point less than the headline step"}
        if hoods_ms:
            line["e2e_hoods_extension"] = {
                "value": world * n * e2e_steps / (hoods_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(n * 8),
                "d2h_bytes_per_step": int(n * NO * 8),
                "note": "ExpertSolver.solve_hoods(f, fi): one value per point crosses PCIe (each rank uploads its slice, NCCL "
                        "all-gathers f), fk = f[hoods] is gathered on the GPU (not the reference's call signature; the "
                        "headline e2e above is)"}
        traffic_file = ROOT / "profiles" / "solve_kernel_traffic.json"
        if traffic_file.exists():
            try:
                tj = json.loads(traffic_file.read_text())
                if tj.get("points") == n:
                    line["roofline"]["traffic"] = tj.get("dram_bytes_per_launch")
                    line["roofline"]["traffic_source"] = tj.get("source")
            except Exception:
                pass
        if world == 1 and not args.no_cpu:
            if all_cpus:
                try:
                    os.sched_setaffinity(0, all_cpus)      # the CPU baseline gets every core of the box
                except Exception:
                    pass
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="points per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling legs (cfg2 1M / cfg3 4M total)")
    ap.add_argument("--probe-variant", default="", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
