#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_loop_gpu.py -m gpu -q -x -k input_pipeline --timeout 600 ) > gpurun_out/test_loop.log 2>&1
tail -40 gpurun_out/test_loop.log
