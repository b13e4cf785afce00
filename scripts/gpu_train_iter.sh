#!/bin/bash
# quick iteration: train-path kernel tests + per-kernel breakdown of the train step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_kernels_gpu.py tests/test_model_gpu.py tests/test_loop_gpu.py -q -x --timeout 300 > gpurun_out/pytest_train.log 2>&1
tail -15 gpurun_out/pytest_train.log
timeout 300 python scripts/profile_step.py mobilenet 32 > gpurun_out/prof_step_mn32.log 2>&1
timeout 300 python scripts/profile_step.py resnet 32 > gpurun_out/prof_step_rn32.log 2>&1
timeout 300 python scripts/profile_step.py mobilenet 4 > gpurun_out/prof_step_mn4.log 2>&1
grep -E "ms/step|encoder|head" gpurun_out/prof_step_*.log
