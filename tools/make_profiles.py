#!/usr/bin/env python
"""Turn the ncu reports of tools/gpu_profile.sh (gpurun_out/*_<tag>.ncu-rep) into the committed summaries under
profiles/:  python tools/make_profiles.py <tag> [<round prefix, default r01>]"""
import csv, io, json, subprocess, sys, shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
import os
OUT = Path(os.environ.get("PROFILES_OUT", str(ROOT / "profiles")))
OUT.mkdir(parents=True, exist_ok=True)
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__cycles_active.avg"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[1], r[2:]


def summarise(name, rep):
    h, units, rows = raw(rep)
    dst = OUT / f"{rnd}_ncu_full_{name}.csv"
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for row in rows:
            d = dict(zip(h, row))
            u = dict(zip(h, units))
            for k in h:
                if k in KEEP or ("issue_stalled" in k and k.endswith("per_issue_active.ratio")):
                    w.writerow([d["Kernel Name"], k, u[k], d[k]])
    lines = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_lines.py"), str(rep), "x", "40"], capture_output=True, text=True).stdout
    (OUT / f"{rnd}_{name}_source_lines.txt").write_text(lines)
    return h, units, rows


G = ROOT / "gpurun_out"
for name, stem in (("solve_kernel", "solve"), ("prepare_reg_kernel", "prepare"), ("interpolate_kernel", "interp"),
                   ("solve_kernel_3d_iter_sens", "solve3d"), ("solve_pack_kernel", "solvepack"),
                   ("solve_kernel_2d_iterative", "solveiter"), ("interpolate_kernel_one_slot", "interp1"),
                   ("fit_direct_kernel", "fitdirect"), ("lu_reg_kernel", "lureg"), ("prepare_reg_kernel_3d", "prepare3d"),
                   ("rescale_kernel", "rescale"), ("sy_thread_kernel", "sythread"), ("prepare_reg_kernel_1d", "prepare1d")):
    rep = G / f"{stem}_{tag}.ncu-rep"
    if not rep.exists():
        print("missing", rep); continue
    h, units, rows = summarise(name, rep)
    if stem == "solve":
        d = dict(zip(h, rows[0])); u = dict(zip(h, units))
        def bytes_of(k):
            v = float(d[k]); return int(v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u[k]])
        rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
        (OUT / "solve_kernel_traffic.json").write_text(json.dumps({
            "kernel": "wlsqm::solve_kernel<1,false,false,true>", "points": 1000000, "dram_bytes_read": rd, "dram_bytes_write": wr,
            "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": 4080000000,
            "source": f"ncu --set full --clock-control none, gpurun_out/solve_{tag}.ncu-rep ({rnd}), launch 1 of 2"}, indent=1) + "\n")
    print("ok", name)
src = G / f"launches_{tag}.csv"
if src.exists():
    keep = [l for l in src.read_text().splitlines() if not l.startswith("==")]
    (OUT / f"{rnd}_launches.csv").write_text("\n".join(keep) + "\n")
for stem in ("configs", "pipeline", "lapack"):
    p = G / f"{stem}_{tag}.jsonl"
    if p.exists():
        shutil.copy(p, OUT / f"{rnd}_{stem}.jsonl")
for stem in ("bench", "bench_ref"):
    p = G / f"{stem}_{tag}.json"
    if p.exists():
        shutil.copy(p, OUT / f"{rnd}_{stem}.json")
p = G / f"oneshot_{tag}.txt"
if p.exists():
    shutil.copy(p, OUT / f"{rnd}_oneshot.txt")
