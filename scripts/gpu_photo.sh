#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_augment_gpu.py tests/test_abi.py -m gpu -q --timeout 600 ) > gpurun_out/test_photo.log 2>&1
tail -30 gpurun_out/test_photo.log
timeout 300 python scripts/bench_augment.py > gpurun_out/bench_augment.log 2>&1; cat gpurun_out/bench_augment.log
