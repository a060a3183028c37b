#!/bin/bash
# final evidence pass of round 2 (after the occupancy changes): randomized sweep, all configs with CPU legs, parity report,
# launch list, full captures of the kernels that changed since the first r02 pass.  usage: bash tools/gpu_profile_r02b.sh <tag>
TAG=${1:-r02b}
mkdir -p gpurun_out
( timeout 1500 python tools/fuzz_sweep.py 300 302 ) > gpurun_out/fuzz_sweep_$TAG.txt 2>&1; grep -c "^ok" gpurun_out/fuzz_sweep_$TAG.txt; tail -1 gpurun_out/fuzz_sweep_$TAG.txt
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
$N -k regex:interpolate_kernel -s 4 -c 1 -o gpurun_out/interp1_$TAG -f python tools/interp_probe.py 2 > gpurun_out/ncu_interp1_$TAG.log 2>&1
$N -k regex:sy_thread_kernel -s 1 -c 1 -o gpurun_out/sythread_$TAG -f python benchmarks/lapack_bench.py --no-cpu > gpurun_out/ncu_sythread_$TAG.log 2>&1
$N -k regex:prepare_reg_kernel -s 1 -c 1 -o gpurun_out/prepare1d_$TAG -f python tools/prep_time.py 2000000 1 3 8 1 > gpurun_out/ncu_prepare1d_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep
PROFILES_OUT=gpurun_out/profiles_$TAG python tools/make_profiles.py $TAG r02 2>&1 | tail -14
for f in gpurun_out/*_$TAG.ncu-rep; do case "$f" in *solve_$TAG.ncu-rep) ;; *) rm -f "$f";; esac; done
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 300 gpurun_out/bench_ref_$TAG.json
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.json; tail -4 gpurun_out/bench_$TAG.err
du -sh gpurun_out
