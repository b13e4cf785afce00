#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_multi.sh "<pytest args>" "<bench args>"'  -> gpurun_out/test.log, bench_1.json
mkdir -p gpurun_out
if [ -n "$1" ]; then
  ( time timeout 900 python -m pytest $1 -m gpu -q -s --timeout 600 ) > gpurun_out/test.log 2>&1
  grep -n "passed\|failed\|error" gpurun_out/test.log | tail -5
fi
if [ -n "$2" ]; then
  ( time timeout 900 python bench.py $2 ) > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
  tail -c 1500 gpurun_out/bench_1.json; echo; tail -5 gpurun_out/bench_1.err
fi
