#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_edge.py -q -k "solution_copy" ) 2>&1 | tail -5
( timeout 600 python tools/fuzz_debug.py 3 1 15; timeout 600 python tools/fuzz_debug.py 2 2 14 ) > gpurun_out/fuzz_debug.txt 2>&1; cat gpurun_out/fuzz_debug.txt | cut -c1-400
