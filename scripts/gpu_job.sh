#!/bin/bash
# generic GPU job: pytest files ($1), extra python scripts ($2, ';'-separated "script args > name"), bench args ($3)
mkdir -p gpurun_out
if [ -n "$1" ]; then
  ( time timeout 900 python -m pytest $1 -m gpu -q -s --timeout 600 ) > gpurun_out/test.log 2>&1
  grep -n "passed\|failed\|error" gpurun_out/test.log | tail -3
fi
if [ -n "$2" ]; then
  IFS=';' read -ra JOBS <<< "$2"
  for j in "${JOBS[@]}"; do
    name=$(echo "$j" | awk '{print $1}' | xargs basename | sed 's/\.py$//')
    ( time timeout 600 python $j ) > gpurun_out/$name.log 2>&1
    tail -40 gpurun_out/$name.log
  done
fi
if [ -n "$3" ]; then
  ( time timeout 900 python bench.py $3 ) > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
  tail -c 1200 gpurun_out/bench_1.json; echo; tail -5 gpurun_out/bench_1.err
fi
