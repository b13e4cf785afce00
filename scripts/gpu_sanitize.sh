#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|smoke ok|Error" gpurun_out/memcheck_smoke.log | head -20
for i in 1 2 3 4 5; do timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1; done
