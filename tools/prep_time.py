#!/usr/bin/env python
"""prepare() timing of the headline cloud (1M fits, 2D order 4, k=30), device-resident inputs.
   python tools/prep_time.py [n] [dim] [order] [k] [knowns]"""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "python-wlsqm_b200")]
import wlsqm_b200 as wlsqm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
order = int(sys.argv[3]) if len(sys.argv) > 3 else 4
k = int(sys.argv[4]) if len(sys.argv) > 4 else 30
kn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
g = torch.Generator(device="cuda").manual_seed(0)
h = 1e-2
xi = h * n ** (1.0 / dim) * torch.rand((n, dim), dtype=torch.float64, device="cuda", generator=g)
xk = xi[:, None, :] + 1.5 * h * (2 * torch.rand((n, k, dim), dtype=torch.float64, device="cuda", generator=g) - 1)
if dim == 1:
    xi, xk = xi[:, 0].contiguous(), xk[:, :, 0].contiguous()
s = wlsqm.ExpertSolver(dim, np.full(n, k, np.int32), np.full(n, order, np.int32), np.full(n, kn, np.int64),
                       np.full(n, wlsqm.WEIGHT_UNIFORM, np.int32), ntasks=1)
ts = []
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.prepare(xi, xk); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("prepare n=%d dim=%d order=%d k=%d knowns=%d: min %.3f ms, median %.3f ms -> %.3g fits/s" % (n, dim, order, k, kn, min(ts[1:]), float(np.median(ts[1:])), n / (min(ts[1:]) * 1e-3)))
