#!/bin/bash
# tests + configs + bench after the substitution / one-shot changes; ncu of prepare 2D + 3D
TAG=${1:-r02e}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
grep -n "^E  " gpurun_out/pytest_gpu_$TAG.log | head -40 | cut -c1-300
timeout 900 python benchmarks/run_configs.py --no-cpu 2>&1 | tee gpurun_out/configs_$TAG.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d.get('impl','b200'), '|', d['config'][:70], '|', d['stage'], d['n'], d['ms'], '%.3g'%d['per_s'], d.get('hbm_frac'))
"
timeout 300 python - <<'PY'
import sys, time
sys.path[:0] = ['.', 'python-wlsqm_b200']
import numpy as np, torch, wlsqm_b200 as wlsqm
# one-shot 3D order 4 (fused kernel) vs prepare + solve, 500k fits, CUDA tensors
n, k = 500_000, 60
g = torch.Generator(device='cuda').manual_seed(0)
xi = 1e-2 * n ** (1 / 3) * torch.rand((n, 3), dtype=torch.float64, device='cuda', generator=g)
xk = xi[:, None, :] + 1.5e-2 * (2 * torch.rand((n, k, 3), dtype=torch.float64, device='cuda', generator=g) - 1)
fk = torch.sin(xk[..., 0]) * torch.cos(xk[..., 1]) * torch.exp(xk[..., 2])
fi = torch.zeros((n, 35), dtype=torch.float64, device='cuda'); fi[:, 0] = torch.sin(xi[:, 0]) * torch.cos(xi[:, 1]) * torch.exp(xi[:, 2])
m = (np.full(n, k, np.int32), np.full(n, 4, np.int32), np.full(n, 1, np.int64), np.full(n, 2, np.int32))
def T(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts)
import os
print('fit_3D_many_parallel o4 k60 500k, fused one-shot kernel: %.2f ms' % T(lambda: wlsqm.fit_3D_many_parallel(xk, fk, m[0], xi, fi, None, 0, m[1], m[2], m[3])))
os.environ['WLSQM_FIT_DIRECT'] = '0'
print('                                  create+prepare+solve  : %.2f ms' % T(lambda: wlsqm.fit_3D_many_parallel(xk, fk, m[0], xi, fi, None, 0, m[1], m[2], m[3])))
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-strong 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('bench ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e'], 'prepare ms', d['prepare']['ms'], d['prepare']['roofline']['frac'], d.get('one_shot_fits'))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/prepare3d_$TAG -f python benchmarks/run_configs.py --no-cpu --only cfg3 --scale 0.25 > gpurun_out/ncu_prepare3d_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/prepare2d_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu --no-strong > gpurun_out/ncu_prepare2d_$TAG.log 2>&1
for r in prepare3d prepare2d; do python tools/ncu_summary.py gpurun_out/${r}_$TAG.ncu-rep > gpurun_out/${r}_${TAG}_summary.txt 2>&1; python tools/ncu_lines.py gpurun_out/${r}_$TAG.ncu-rep x 60 > gpurun_out/${r}_${TAG}_lines.txt 2>&1; done
head -22 gpurun_out/prepare3d_${TAG}_summary.txt; head -22 gpurun_out/prepare2d_${TAG}_summary.txt
