#!/bin/bash
# tests + configs timing after the prepare changes
TAG=${1:-r02c}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -8 gpurun_out/pytest_gpu_$TAG.log
grep -n "^E  " gpurun_out/pytest_gpu_$TAG.log | head -40 | cut -c1-300
timeout 900 python benchmarks/run_configs.py --no-cpu 2>&1 | tee gpurun_out/configs_$TAG.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['config'][:70], '|', d['stage'], d['n'], d['ms'], '%.3g'%d['per_s'], d.get('hbm_frac'))
"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-strong 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('bench ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'prepare ms', d['prepare']['ms'], d['prepare']['roofline']['frac'], d.get('one_shot_fits'))"
