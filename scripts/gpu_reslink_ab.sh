#!/bin/bash
# A/B of the shortcut-gradient link (kEpiRawRes): tests, then the headline train leg with PP_RES_LINK=1 / 0
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_encoder_convs_gpu.py tests/test_train_parity_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 600 ) > gpurun_out/test.log 2>&1
grep -n "passed\|failed\|error" gpurun_out/test.log | tail -3
for v in 1 0 1 0; do
  PP_RES_LINK=$v timeout 600 python bench.py --no-extras --no-query --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_link$v.json 2> gpurun_out/bench_link$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_link$v.json").read().strip().splitlines()[-1])
print("PP_RES_LINK=$v", d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done
