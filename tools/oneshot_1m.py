"""one-shot fit_2D_many_parallel on 1M points, 2D order 4, k = 30 (CUDA tensors): wall time and GPU time per call"""
import sys, time
sys.path[:0] = ["/root/repo", "/root/repo/python-wlsqm_b200"]
import numpy as np, torch, wlsqm_b200 as wlsqm
n, k = 1_000_000, 30
g = torch.Generator(device="cuda").manual_seed(0)
xi = 10 * torch.rand((n, 2), dtype=torch.float64, device="cuda", generator=g)
xk = xi[:, None, :] + 0.015 * (2 * torch.rand((n, k, 2), dtype=torch.float64, device="cuda", generator=g) - 1)
fk = torch.sin(xk[..., 0]) * torch.cos(xk[..., 1])
fi = torch.zeros((n, 15), dtype=torch.float64, device="cuda")
m = (np.full(n, k, np.int32), np.full(n, 4, np.int32), np.zeros(n, np.int64), np.full(n, 1, np.int32))
for rep in range(5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    wlsqm.fit_2D_many_parallel(xk, fk, m[0], xi, fi, None, 0, m[1], m[2], m[3])
    e1.record()
    torch.cuda.synchronize()
    print("wall %.3f ms, gpu-side %.3f ms" % (1e3 * (time.perf_counter() - t0), e0.elapsed_time(e1)))
