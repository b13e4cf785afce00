#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_query_selector_gpu.py tests/test_loop_gpu.py -m gpu -q --timeout 500 ) > gpurun_out/test_q.log 2>&1
grep -n "passed\|failed" gpurun_out/test_q.log | tail -2
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print("train", d["value"], "query_rn50_mpix_s", d["config"]["query_rn50_mpix_s"], d["config"]["query_rn50_img_s"])
PY
