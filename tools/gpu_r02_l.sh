#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "random_knowns or every_size" ) 2>&1 | tail -30 | cut -c1-300
( timeout 2400 python tools/fuzz_sweep.py 200 204 ) > gpurun_out/fuzz_sweep.txt 2>&1; grep -c "^ok" gpurun_out/fuzz_sweep.txt; grep -A14 "^FAIL" gpurun_out/fuzz_sweep.txt | cut -c1-400 | head -150; tail -1 gpurun_out/fuzz_sweep.txt
