#!/bin/bash
# quick GPU check: gpu tests + headline bench + secondary configs; usage: bash tools/gpu_quick.sh <tag> [configs-filter]
TAG=${1:-x}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log; grep -n "^E " gpurun_out/pytest_gpu_$TAG.log | head -20
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('bench ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'prepare ms', d['prepare']['ms'])"
timeout 1200 python benchmarks/run_configs.py ${2:+--only $2} 2>&1 | tee gpurun_out/configs_$TAG.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['config'][:60], d['stage'], d['n'], d['ms'], '%.3g'%d['per_s'], d.get('hbm_frac'))
"
