#!/bin/bash
# validation pass: gpu tests, smoke(), both bench arms, memcheck over the kernels' tests
TAG=${1:-r02final}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
grep -n "^E  " gpurun_out/pytest_gpu_$TAG.log | head -20 | cut -c1-300
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_$TAG.log 2>&1; tail -3 gpurun_out/smoke_$TAG.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 600 gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; python - <<PY
import json
l=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print({k:l.get(k) for k in ("value","ms_per_step","gpu_launches","clocks")}); print(l["e2e"]["value"], l["roofline"], l["cpu_baseline"]["value"])
PY
( time timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_edge.py tests/test_gpu_variants.py tests/test_gpu_golden.py tests/test_gpu_parity.py -k "not fullsize and not large_pageable and not side_streams and not every_size and not ab_switches" ) > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; tail -8 gpurun_out/r02_sanitizer_memcheck.txt
