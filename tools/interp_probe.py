"""one-slot interpolate (the reference's interface) launch-shape probe: python tools/interp_probe.py DIM
(run once per WLSQM_INTERP_MINB / WLSQM_INTERP_Q setting: the library reads them once per process)"""
import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/python-wlsqm_b200"]
import numpy as np, torch, wlsqm_b200 as wlsqm
dim = int(sys.argv[1])
nm, per = (1_000_000, 16) if dim < 3 else (500_000, 16)
no = wlsqm.number_of_dofs(dim, 4)
g = torch.Generator(device="cuda").manual_seed(0)
xi = torch.rand((nm, dim), dtype=torch.float64, device="cuda", generator=g)
fi = torch.rand((nm, no), dtype=torch.float64, device="cuda", generator=g)
I = torch.arange(nm, device="cuda").repeat_interleave(per)
x = xi[I] + 1e-3 * torch.rand((nm * per, dim), dtype=torch.float64, device="cuda", generator=g)
if dim == 1:
    xi, x = xi[:, 0].contiguous(), x[:, 0].contiguous()
z = np.zeros(nm, np.int32)
s = wlsqm.ExpertSolver(dim, z + 1, z + 4, np.zeros(nm, np.int64), z + 1)
# models straight into the solver: prepare on a trivial geometry, then solve() is bypassed by loading fi through the guest-free path
xk = (xi.reshape(nm, 1, -1) + 0.01).reshape((nm, 1) if dim == 1 else (nm, 1, dim))
s.prepare(xi, xk)
s.solve(torch.zeros((nm, 1), dtype=torch.float64, device="cuda"), fi.clone())
s.prep_interpolate(search='gpu')
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
nq = nm * per
byt = 8.0 * (dim + 2) + (8.0 * no + 8 * dim) / per
for d in (0, no - 1):
    ms = t(lambda: s.interpolate(x, diff=d, I=I))
    print("%dD MINB=%s Q=%s diff %d: %.4f ms per %dM queries (%.2f of 6526 GB/s at %.1f B/query)" % (
        dim, os.environ.get("WLSQM_INTERP_MINB", "default"), os.environ.get("WLSQM_INTERP_Q", "default"), d, ms, nq // 1_000_000,
        byt * nq / ms / 1e6 / 6526, byt), flush=True)
