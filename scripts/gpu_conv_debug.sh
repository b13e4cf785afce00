#!/bin/bash
# conv kernel bring-up: bounded by `timeout` so a deadlocked kernel cannot hang the box
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -q -x --timeout 120 2>&1 | tail -30 > gpurun_out/conv_pytest.log
cat gpurun_out/conv_pytest.log
for v in 0 1 2 3; do PP_SCORE_VARIANT=$v timeout 300 python scripts/bench_score_variants.py; done 2>&1 | tee gpurun_out/score_variants.log | tail -8
