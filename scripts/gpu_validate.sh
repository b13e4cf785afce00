#!/bin/bash
# One gpurun call: GPU tests, smoke, both bench arms, per-kernel breakdown of the train step (torch.profiler).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --durations=15 ) > gpurun_out/pytest_full.log 2>&1
tail -30 gpurun_out/pytest_full.log > gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python scripts/profile_step.py mobilenet 32 > gpurun_out/prof_step_mn32.log 2>&1
timeout 300 python scripts/profile_step.py resnet 32 > gpurun_out/prof_step_rn32.log 2>&1
timeout 300 python scripts/profile_step.py mobilenet 4 > gpurun_out/prof_step_mn4.log 2>&1
tail -8 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
