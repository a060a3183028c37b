#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "random_knowns" ) 2>&1 | tail -15 | cut -c1-300
( timeout 1500 python tools/fuzz_sweep.py 100 106 ) > gpurun_out/fuzz_sweep.txt 2>&1; grep -c "^ok" gpurun_out/fuzz_sweep.txt; grep -A12 "^FAIL" gpurun_out/fuzz_sweep.txt | cut -c1-400 | head -120; tail -1 gpurun_out/fuzz_sweep.txt
