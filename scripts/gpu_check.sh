#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list and one full capture of the scoring kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40 > gpurun_out/pytest.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "${1:-}" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-batch 16 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:acq_score_vec -s 3 -c 1 -o gpurun_out/prof_score -f \
    python bench.py --steps 2 --warmup 3 --e2e-batch 16 --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
fi
tail -5 gpurun_out/pytest.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
