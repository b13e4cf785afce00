#!/bin/bash
# quick GPU check of selected test files (bounded so a deadlocked kernel cannot hang the box)
mkdir -p gpurun_out
timeout 900 python -m pytest "$@" -q --timeout 240 --durations=8 -s > gpurun_out/quick_full.log 2>&1
grep -E "rel |cos |agree|passed|failed|^FAILED|^E  " gpurun_out/quick_full.log | head -${LINES_MAX:-80}
