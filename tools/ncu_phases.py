#!/usr/bin/env python
"""Instruction / stall-sample share per source-line range of one file in an ncu report:
   python tools/ncu_phases.py rep.ncu-rep file.cu name:lo-hi name:lo-hi ..."""
import csv, io, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
ranges = []
for a in sys.argv[3:]:
    n, r = a.split(":"); lo, hi = r.split("-"); ranges.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, cur = None, None
acc = {n: [0, 0] for n, _, _ in ranges}; other = [0, 0]; tot = [0, 0]
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]
    elif r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:]))
        try: inst = int(d["Instructions Executed"]); smp = int(d["# Samples"] or 0)
        except ValueError: continue
        tot[0] += inst; tot[1] += smp
        hit = False
        if cur == fname:
            ln = int(r[0])
            for n, lo, hi in ranges:
                if lo <= ln <= hi: acc[n][0] += inst; acc[n][1] += smp; hit = True; break
        if not hit: other[0] += inst; other[1] += smp
print("total inst %d samples %d" % tuple(tot))
for n, _, _ in ranges: print("%-12s %5.1f%% inst %5.1f%% samples" % (n, 100*acc[n][0]/tot[0], 100*acc[n][1]/tot[1]))
print("%-12s %5.1f%% inst %5.1f%% samples" % ("other", 100*other[0]/tot[0], 100*other[1]/tot[1]))
