#!/bin/bash
# reference's own tests against the drop-in (Cython shim) + compute-sanitizer over the new kernels + gpu tests
TAG=${1:-r02f}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
grep -n "^E  " gpurun_out/pytest_gpu_$TAG.log | head -20 | cut -c1-300
( timeout 600 python tools/reference_tests_against_b200.py --run -v ) > gpurun_out/r02_reference_tests_against_b200.txt 2>&1; tail -15 gpurun_out/r02_reference_tests_against_b200.txt
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_edge.py tests/test_gpu_variants.py tests/test_gpu_golden.py "tests/test_gpu_parity.py" -k "not fullsize and not large_pageable and not side_streams" ) > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; tail -8 gpurun_out/r02_sanitizer_memcheck.txt
( time timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_variants.py -k "one_shot or packed or prepare_variants" tests/test_gpu_parity.py::test_batched_scalers_vs_reference_golden tests/test_gpu_parity.py::test_single_system_drivers_vs_lapack ) > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; tail -8 gpurun_out/r02_sanitizer_racecheck.txt
