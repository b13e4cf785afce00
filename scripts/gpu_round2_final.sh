#!/bin/bash
# Round-2 evidence run on ONE GPU:  gpurun --timeout 2400 -- 'bash scripts/gpu_round2_final.sh'
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --durations=8 ) > gpurun_out/pytest_full.log 2>&1
tail -14 gpurun_out/pytest_full.log > gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
( time timeout 900 python bench.py --steps 20 --warmup 5 --sweep ) > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
# per-layer / per-kernel breakdowns
timeout 300 python scripts/bench_enc_train_layers.py 32 > gpurun_out/bench_enc_train_layers.log 2>&1
timeout 300 python scripts/bench_wgrad.py > gpurun_out/bench_wgrad.log 2>&1
timeout 300 python scripts/profile_step.py resnet 32 > gpurun_out/prof_step_rn32.log 2>&1
timeout 300 python scripts/profile_step.py resnet 4 > gpurun_out/prof_step_rn4.log 2>&1
timeout 300 python scripts/profile_step.py mobilenet 32 > gpurun_out/prof_step_mn32.log 2>&1
# ncu launch list of the bench command (headline train leg + acquisition leg; graph replays are profiled per node)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-query --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# acquisition pipeline: launch list with DRAM bytes, then full captures of the dominant kernels
bash scripts/gpu_ncu_list.sh acq "scripts/acq_step.py 256 3" 200 > gpurun_out/acq_list.log 2>&1
bash scripts/gpu_ncu.sh ncu_acq "scripts/acq_step.py 256 1" "acq_score|select_l0|pick_ranks_fast" 3 > /dev/null 2>&1
bash scripts/gpu_ncu.sh ncu_l1c3 "scripts/profile_conv_layer.py 64 256 1 1 64 128 32" "conv_igemm|wgrad|nvjet|cutlass" 6 > /dev/null 2>&1
bash scripts/gpu_ncu.sh ncu_l4c2 "scripts/profile_conv_layer.py 512 512 3 4 32 64 32" "conv_igemm|wgrad|nvjet|cutlass" 6 > /dev/null 2>&1
tail -4 gpurun_out/pytest.log; tail -1 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("bench_ref", "bench_1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    except Exception as e:
        print(f, "FAILED", e)
PY
