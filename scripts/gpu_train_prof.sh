#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_train_kernels_gpu.py tests/test_loop_gpu.py -q --timeout 300 2>&1 | tail -25
timeout 300 python scripts/profile_step.py mobilenet 32 2>&1 | grep -v Warning | grep -vE "^    " | tail -8
timeout 300 python scripts/profile_step.py resnet 32 2>&1 | grep -v Warning | grep -vE "^    "| tail -8
