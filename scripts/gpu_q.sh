#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/debug_smoke.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_acq_gpu.py tests/test_query_selector_gpu.py -q -x --timeout 180 2>&1 | tail -5
timeout 300 python bench.py --no-train --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; cat gpurun_out/bench_q.json | python -c "
import json,sys; d=json.load(sys.stdin); print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'score_ms',d['roofline']['kernel_ms'],'e2e',d['e2e']['value'])"; tail -3 gpurun_out/bench_q.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_q.csv python bench.py --steps 2 --warmup 3 --e2e-batch 16 --no-cpu-baseline --no-train > /dev/null 2>&1
