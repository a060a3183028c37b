import sys
sys.path[:0] = ["/root/repo", "/root/repo/python-wlsqm_b200"]
import numpy as np, torch
from wlsqm_b200.utils import lapackdrivers as ld
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
for n in (3, 6, 8, 10, 12, 15, 16):
    nlhs = 1_000_000 if n <= 8 else 500_000
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn((nlhs, n, n), dtype=torch.float64, device="cuda", generator=g)
    A = 0.5 * (A + A.transpose(1, 2))
    b = torch.randn((nlhs, n), dtype=torch.float64, device="cuda", generator=g)
    def run():
        A2 = A.clone().permute(2, 1, 0); b2 = b.clone().t()
        return A2, b2
    A2, b2 = run()
    # time the driver alone on fresh copies (the clone is outside the events)
    ts = []
    for _ in range(5):
        A2, b2 = run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ld.msymmetric(A2, b2); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[1:]))
    print("msymmetric n=%d nlhs=%d: %.3f ms = %.3f ns per system" % (n, nlhs, ms, ms * 1e6 / nlhs), flush=True)
