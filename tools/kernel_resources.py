#!/usr/bin/env python
"""Static resources of every kernel in libwlsqm_b200.so -> profiles/<tag>_kernel_resources.txt
(registers, stack, static shared memory, local memory from `cuobjdump --dump-resource-usage`; SASS instruction
count from `cuobjdump -sass`).  Needs no GPU.  Usage: python tools/kernel_resources.py [tag]"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
so = str(ROOT / "python-wlsqm_b200" / "wlsqm_b200" / "libwlsqm_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True, check=True).stdout
rows, name = [], None
for line in out.splitlines():
    line = line.strip()
    if line.startswith("Function "):
        name = line[len("Function "):].rstrip(":")
    elif line.startswith("REG:") and name:
        d = dict(t.split(":") for t in line.split() if ":" in t and not t.startswith("CONSTANT"))
        rows.append((name, int(d["REG"]), int(d["STACK"]), int(d["SHARED"]), int(d["LOCAL"])))
        name = None
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
counts, cur = {}, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = 0
    elif cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
        counts[cur] += 1
dst = ROOT / "profiles" / ("%s_kernel_resources.txt" % tag)
with open(dst, "w") as f:
    f.write("# static resources of every kernel in libwlsqm_b200.so (cuobjdump --dump-resource-usage / -sass), sm_100a\n")
    f.write("# %-100s %5s %6s %7s %6s %8s\n" % ("kernel", "regs", "stack", "static", "local", "SASS"))
    for (mangled, reg, stack, sh, loc), nm in sorted(zip(rows, names), key=lambda t: t[1]):
        nm = nm.replace("wlsqm::", "").replace("(anonymous namespace)::", "")
        nm = re.sub(r"\(.*\)$", "", nm)
        f.write("%-102s %5d %6d %7d %6d %8d\n" % (nm[:102], reg, stack, sh, loc, counts.get(mangled, -1)))
print("wrote", dst)
