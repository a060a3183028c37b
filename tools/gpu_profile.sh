#!/bin/bash
# GPU-box script: tests, bench, ncu launch list and full captures of the kernels of the path.
# usage (under gpurun): bash tools/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
grep -n "^E " gpurun_out/pytest_gpu_$TAG.log | head -30
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; tail -c 600 gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_$TAG.log 2>&1
grep -c . gpurun_out/launches_$TAG.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 3 -c 2 -o gpurun_out/solve_$TAG -f python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_full_solve_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/prepare_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_prepare_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interpolate_kernel -s 2 -c 2 -o gpurun_out/interp_$TAG -f python benchmarks/run_configs.py --only cfg5 > gpurun_out/ncu_interp_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 2 -c 1 -o gpurun_out/solve3d_$TAG -f python benchmarks/run_configs.py --only cfg3 --scale 0.25 > gpurun_out/ncu_solve3d_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_pack_kernel -s 2 -c 2 -o gpurun_out/solvepack_$TAG -f python benchmarks/run_configs.py --only cfg4 > gpurun_out/ncu_solvepack_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 12 -c 1 -o gpurun_out/solveiter_$TAG -f python benchmarks/run_configs.py --only cfg2v > gpurun_out/ncu_solveiter_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interpolate_kernel -s 9 -c 1 -o gpurun_out/interp1_$TAG -f python benchmarks/run_configs.py --only cfg5 > gpurun_out/ncu_interp1_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/fitdirect_$TAG -f python tools/oneshot_1m.py > gpurun_out/ncu_fitdirect_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lu_reg_kernel -s 8 -c 1 -o gpurun_out/lureg_$TAG -f python benchmarks/lapack_bench.py --no-cpu > gpurun_out/ncu_lureg_$TAG.log 2>&1
timeout 600 python benchmarks/lapack_bench.py > gpurun_out/lapack_$TAG.jsonl 2>&1
timeout 300 python tools/oneshot_probe.py > gpurun_out/oneshot_$TAG.txt 2>&1
timeout 1200 python benchmarks/run_configs.py > gpurun_out/configs_$TAG.jsonl 2>&1; tail -25 gpurun_out/configs_$TAG.jsonl | cut -c1-220
timeout 600 python benchmarks/pipeline.py --host-tree > gpurun_out/pipeline_$TAG.jsonl 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep
# the reports together exceed what gpurun copies back (64 MiB): summarise them here, keep only the headline report
PROFILES_OUT=gpurun_out/profiles_$TAG python tools/make_profiles.py $TAG r01 2>&1 | tail -15
for f in gpurun_out/*_$TAG.ncu-rep; do case "$f" in *solve_$TAG.ncu-rep) ;; *) rm -f "$f";; esac; done
du -sh gpurun_out
