#!/bin/bash
# quick iteration on the query path: parity tests + query-only bench + ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_acq_gpu.py tests/test_query_selector_gpu.py -q -x --timeout 300 2>&1 | tail -4
timeout 600 python bench.py --no-train --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python -c "
import json; d=json.load(open('gpurun_out/bench_q.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'score_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_q.csv python bench.py --steps 2 --warmup 3 --e2e-batch 16 --no-cpu-baseline --no-train > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_q.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[-12:]:
    print(r[4][:60], r[-1])
PY
