#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_test.sh tests/test_x.py [-k expr]'  -> gpurun_out/test.log
mkdir -p gpurun_out
( time timeout 800 python -m pytest "$@" -m gpu -q -s --timeout 600 ) > gpurun_out/test.log 2>&1
grep -v "^$" gpurun_out/test.log | tail -60
