#!/bin/bash
# One gpurun call producing everything profiles/ needs: tests, smoke, both bench arms, ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --durations=8 ) > gpurun_out/pytest_full.log 2>&1
tail -14 gpurun_out/pytest_full.log > gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python scripts/bench_dw.py 32 2>&1 | grep -v Warn > gpurun_out/bench_dw.log
timeout 300 python scripts/bench_enc_layers.py 32 2>&1 | grep -v Warn > gpurun_out/bench_enc_layers.log
timeout 300 python scripts/bench_eval_fwd.py 2>&1 | grep -v Warn > gpurun_out/bench_eval_fwd.log
timeout 300 python scripts/profile_step.py mobilenet 32 > gpurun_out/prof_step_mn32.log 2>&1
timeout 300 python scripts/profile_step.py resnet 32 > gpurun_out/prof_step_rn32.log 2>&1
if [ "${1:-}" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-batch 16 --no-cpu-baseline --no-train > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"acq_score_vec|select_l0|select_rest|pick_ranks|pick_bucket0" -s 12 -c 6 -o gpurun_out/prof_topk -f \
    python bench.py --steps 2 --warmup 3 --e2e-batch 16 --no-cpu-baseline --no-train > gpurun_out/bench_under_ncu_full.log 2>&1
if [ "${1:-}" != "qncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm|wgrad_kernel" -c 14 -o gpurun_out/prof_conv -f \
    python scripts/profile_conv.py > gpurun_out/profile_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dwconv|bn_fwd_fused|bn_bwd_fused" -s 8 -c 10 -o gpurun_out/prof_dwbn -f \
    python scripts/bench_dw.py 32 > gpurun_out/profile_dwbn.log 2>&1
fi
fi
tail -4 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; head -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
