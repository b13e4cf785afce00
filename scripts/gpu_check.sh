#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list and full captures of the top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --durations=15 > gpurun_out/pytest_full.log 2>&1
tail -30 gpurun_out/pytest_full.log > gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "${1:-}" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-batch 16 --no-cpu-baseline --no-train > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"acq_score_vec|select_level|select_rest|pick_ranks" -s 9 -c 4 -o gpurun_out/prof_topk -f \
    python bench.py --steps 2 --warmup 3 --e2e-batch 16 --no-cpu-baseline --no-train > gpurun_out/bench_under_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm|wgrad" -c 6 -o gpurun_out/prof_conv -f \
    python scripts/profile_conv.py > gpurun_out/profile_conv.log 2>&1
fi
tail -8 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
