#!/bin/bash
# tests (all failures listed) + parity report
TAG=${1:-r02b}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -15 gpurun_out/pytest_gpu_$TAG.log
grep -n "^E  " gpurun_out/pytest_gpu_$TAG.log | head -60 | cut -c1-300
( time timeout 1200 python tools/parity_report.py ) > gpurun_out/r02_parity_report.txt 2> gpurun_out/parity_report_$TAG.err; tail -5 gpurun_out/parity_report_$TAG.err; grep -c . gpurun_out/r02_parity_report.txt
