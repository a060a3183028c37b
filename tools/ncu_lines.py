#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo and --import-source on):
   python tools/ncu_lines.py gpurun_out/x.ncu-rep [kernel-regex] [top N]
Prints the source lines with the most executed warp instructions and their stall samples."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, cur_file, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:]))
        try:
            inst = int(d["Instructions Executed"])
        except ValueError:
            continue
        lines.append((inst, int(d["# Samples"] or 0), cur_file, r[0], r[1].strip(), d))
tot = sum(l[0] for l in lines) or 1
tots = sum(l[1] for l in lines) or 1
print(f"total warp instructions {tot:,}; samples {tots:,}")
for inst, smp, f, ln, src, d in sorted(lines, key=lambda l: -l[0])[:top]:
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")}
    ts = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*inst/tot:5.1f}% inst {100*smp/tots:5.1f}% smp  {f}:{ln:>4}  {src[:90]:90s} {ts}")
