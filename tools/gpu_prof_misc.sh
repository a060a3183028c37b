#!/bin/bash
# ncu full captures of the secondary kernels; usage: bash tools/gpu_prof_misc.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interpolate_kernel -s 2 -c 2 -o gpurun_out/interp_$TAG -f python benchmarks/run_configs.py --only cfg5 > gpurun_out/ncu_interp_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 12 -c 1 -o gpurun_out/solveiter_$TAG -f python benchmarks/run_configs.py --only cfg2v > gpurun_out/ncu_solveiter_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 2 -c 1 -o gpurun_out/solve3d_$TAG -f python benchmarks/run_configs.py --only cfg3 --scale 0.25 > gpurun_out/ncu_solve3d_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep
