#!/usr/bin/env python
"""Key raw metrics of every kernel in an ncu report: python tools/ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h = r[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "lts__t_bytes.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
stall = [k for k in h if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
for row in r[2:]:
    d = dict(zip(h, row))
    print("=" * 100)
    for k in want:
        if k in d: print("%-70s %s" % (k, d[k]))
    ss = sorted(((float(d[k] or 0), k.split("issue_stalled_")[1].split("_per_issue")[0]) for k in stall), reverse=True)[:6]
    print("stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in ss))
