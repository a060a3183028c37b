#!/bin/bash
# tests + configs (with CPU figures) + one-shot probe + ncu of the new prepare kernel
TAG=${1:-r02d}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -8 gpurun_out/pytest_gpu_$TAG.log
grep -n "^E  " gpurun_out/pytest_gpu_$TAG.log | head -40 | cut -c1-300
( time timeout 1500 python benchmarks/run_configs.py ) 2>&1 | tee gpurun_out/configs_$TAG.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d.get('impl','b200'), '|', d['config'][:70], '|', d['stage'], d['n'], d['ms'], '%.3g'%d['per_s'], d.get('hbm_frac'), d.get('ntasks'))
"
timeout 300 python tools/oneshot_probe.py > gpurun_out/oneshot_$TAG.txt 2>&1; tail -12 gpurun_out/oneshot_$TAG.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/prepare2d_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu --no-strong > gpurun_out/ncu_prepare2d_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prepare2d_$TAG.ncu-rep > gpurun_out/prepare2d_${TAG}_summary.txt 2>&1; python tools/ncu_lines.py gpurun_out/prepare2d_$TAG.ncu-rep x 70 > gpurun_out/prepare2d_${TAG}_lines.txt 2>&1
cat gpurun_out/prepare2d_${TAG}_summary.txt | head -30
