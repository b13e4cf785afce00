#!/bin/bash
# one-launch Adam (pp_adam_step_multi): tests, then the train legs with PP_ADAM=ours / torch
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_loop_gpu.py tests/test_conv_gpu.py tests/test_encoder_convs_gpu.py tests/test_train_parity_gpu.py -m gpu -q --timeout 600 ) > gpurun_out/test.log 2>&1
grep -n "passed\|failed\|error" gpurun_out/test.log | tail -3
for v in ours; do
  PP_ADAM=$v timeout 600 python bench.py --no-query --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_adam_$v.json 2> gpurun_out/bench_adam_$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_adam_$v.json").read().strip().splitlines()[-1])
c = d["config"]
print("PP_ADAM=$v", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "rn50 b4", c.get("train_rn50_b4_img_s"), "mnv2 b32", c.get("train_mnv2_b32_img_s"), "b4", c.get("train_mnv2_b4_img_s"))
PY
done
timeout 300 python scripts/profile_step.py resnet 32 > gpurun_out/prof_step_rn32.log 2>&1
grep -n "adam\|Adam\|pack_weight" gpurun_out/prof_step_rn32.log | head -8
