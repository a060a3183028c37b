#!/bin/bash
# multi-GPU call (gpurun --gpus 8): host<->device ceiling sweep, bench at N = 8 and 2, sharded-vs-unsharded check
TAG=${1:-r02m}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1; nvidia-smi -L | wc -l
( time timeout 600 python benchmarks/pcie_ceiling.py --steps 10 --out gpurun_out/r02_pcie_ceiling.json ) 2>&1 | tail -16 | cut -c1-400
NG=$(nvidia-smi -L | wc -l)
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 ) > gpurun_out/bench_${NG}gpu_$TAG.json 2> gpurun_out/bench_${NG}gpu_$TAG.err; tail -c 5000 gpurun_out/bench_${NG}gpu_$TAG.json; tail -5 gpurun_out/bench_${NG}gpu_$TAG.err | cut -c1-300
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 benchmarks/multi_gpu_check.py ) > gpurun_out/multi_gpu_check_${NG}gpu_$TAG.json 2>&1; tail -3 gpurun_out/multi_gpu_check_${NG}gpu_$TAG.json | cut -c1-900
if [ "$NG" -gt 2 ]; then
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/bench_2gpu_$TAG.json 2> gpurun_out/bench_2gpu_$TAG.err; tail -c 3000 gpurun_out/bench_2gpu_$TAG.json; tail -3 gpurun_out/bench_2gpu_$TAG.err | cut -c1-300
fi
