#!/bin/bash
# round 2, first GPU call: tests, bench (both arms), parity report, ncu captures of the prepare kernels, PCIe ceiling (N=1)
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
grep -n "^E " gpurun_out/pytest_gpu_$TAG.log | head -30
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 6000 gpurun_out/bench_$TAG.json; tail -8 gpurun_out/bench_$TAG.err
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 1800 gpurun_out/bench_ref_$TAG.json; tail -4 gpurun_out/bench_ref_$TAG.err
( time timeout 1200 python tools/parity_report.py ) > gpurun_out/r02_parity_report.txt 2> gpurun_out/parity_report_$TAG.err; tail -40 gpurun_out/r02_parity_report.txt; tail -5 gpurun_out/parity_report_$TAG.err
timeout 600 python benchmarks/pcie_ceiling.py --out gpurun_out/pcie_ceiling_n1_$TAG.json 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/prepare3d_$TAG -f python benchmarks/run_configs.py --only cfg3 --scale 0.25 > gpurun_out/ncu_prepare3d_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/prepare2d_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu --no-strong > gpurun_out/ncu_prepare2d_$TAG.log 2>&1
for r in prepare3d prepare2d; do python tools/ncu_summary.py gpurun_out/${r}_$TAG.ncu-rep > gpurun_out/${r}_${TAG}_summary.txt 2>&1; python tools/ncu_lines.py gpurun_out/${r}_$TAG.ncu-rep x 60 > gpurun_out/${r}_${TAG}_lines.txt 2>&1; done
cat gpurun_out/prepare3d_${TAG}_summary.txt | head -40
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
