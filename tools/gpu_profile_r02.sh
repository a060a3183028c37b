#!/bin/bash
# GPU-box script of round 2: tests, both bench arms, all configs with the reference's CPU figures, ncu launch list and full
# captures of the kernels that changed this round.  usage (under gpurun): bash tools/gpu_profile_r02.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
grep -n "^E  " gpurun_out/pytest_gpu_$TAG.log | head -30 | cut -c1-300
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 2500 gpurun_out/bench_$TAG.json; tail -4 gpurun_out/bench_$TAG.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 1200 gpurun_out/bench_ref_$TAG.json; tail -4 gpurun_out/bench_ref_$TAG.err
( time timeout 1500 python benchmarks/run_configs.py ) > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; python - <<PY
import json
for l in open('gpurun_out/configs_$TAG.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d.get('impl','b200'), '|', d['config'][:64], '|', d['stage'], d['n'], d['ms'], '%.3g'%d['per_s'], d.get('hbm_frac'))
PY
timeout 900 python tools/parity_report.py > gpurun_out/r02_parity_report.txt 2> gpurun_out/parity_report_$TAG.err; grep -c "yes" gpurun_out/r02_parity_report.txt; grep -c " NO" gpurun_out/r02_parity_report.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-strong > gpurun_out/bench_under_ncu_$TAG.log 2>&1
grep -c . gpurun_out/launches_$TAG.csv
N="timeout 600 ncu --set full --clock-control none --import-source on"
$N -k regex:solve_kernel -s 3 -c 2 -o gpurun_out/solve_$TAG -f python bench.py --steps 4 --warmup 3 --no-cpu --no-strong > gpurun_out/ncu_solve_$TAG.log 2>&1
$N -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/prepare_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu --no-strong > gpurun_out/ncu_prepare_$TAG.log 2>&1
$N -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/prepare3d_$TAG -f python benchmarks/run_configs.py --no-cpu --only cfg3 --scale 0.25 > gpurun_out/ncu_prepare3d_$TAG.log 2>&1
$N -k regex:solve_kernel -s 2 -c 1 -o gpurun_out/solve3d_$TAG -f python benchmarks/run_configs.py --no-cpu --only cfg3 --scale 0.25 > gpurun_out/ncu_solve3d_$TAG.log 2>&1
$N -k regex:solve_kernel -s 12 -c 1 -o gpurun_out/solveiter_$TAG -f python benchmarks/run_configs.py --no-cpu --only cfg2v > gpurun_out/ncu_solveiter_$TAG.log 2>&1
$N -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/fitdirect_$TAG -f python tools/oneshot_1m.py > gpurun_out/ncu_fitdirect_$TAG.log 2>&1
$N -k regex:rescale_kernel -s 1 -c 1 -o gpurun_out/rescale_$TAG -f python benchmarks/lapack_bench.py --no-cpu --scalers > gpurun_out/ncu_rescale_$TAG.log 2>&1
timeout 600 python benchmarks/lapack_bench.py --scalers > gpurun_out/lapack_$TAG.jsonl 2>&1; tail -8 gpurun_out/lapack_$TAG.jsonl | cut -c1-250
ls -la gpurun_out/*_$TAG.ncu-rep
PROFILES_OUT=gpurun_out/profiles_$TAG python tools/make_profiles.py $TAG r02 2>&1 | tail -12
for f in gpurun_out/*_$TAG.ncu-rep; do case "$f" in *solve_$TAG.ncu-rep) ;; *) rm -f "$f";; esac; done
du -sh gpurun_out
