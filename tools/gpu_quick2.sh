#!/bin/bash
# quick check: gpu tests + configs timing (no CPU legs) + bench prepare
TAG=${1:-q}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
grep -n "^E  " gpurun_out/pytest_gpu_$TAG.log | head -20 | cut -c1-300
timeout 900 python benchmarks/run_configs.py --no-cpu ${2:+--only $2} 2>&1 | tee gpurun_out/configs_$TAG.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d.get('impl','b200'), '|', d['config'][:70], '|', d['stage'], d['n'], d['ms'], '%.3g'%d['per_s'], d.get('hbm_frac'))
"
