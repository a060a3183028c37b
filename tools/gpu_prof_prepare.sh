#!/bin/bash
# ncu full capture of the prepare kernel (2D order 4 via bench.py); usage: bash tools/gpu_prof_prepare.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/prepare_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_prepare_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_prepare_$TAG.log
ls -la gpurun_out/*.ncu-rep
