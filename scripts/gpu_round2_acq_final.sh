#!/bin/bash
# Round-2 closing evidence run on ONE GPU (lean: the acquisition kernels changed last):  gpurun --timeout 700 -- 'bash scripts/gpu_round2_acq_final.sh'
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
tail -1 gpurun_out/smoke.log
( timeout 300 python -m pytest tests/test_acq_gpu.py tests/test_query_selector_gpu.py tests/test_loop_gpu.py -q ) > gpurun_out/pytest_acq.log 2>&1
tail -1 gpurun_out/pytest_acq.log
# acquisition pipeline: launch list with DRAM bytes, then full captures of its kernels
bash scripts/gpu_ncu_list.sh acq "scripts/acq_step.py 256 3" 200 > gpurun_out/acq_list.log 2>&1
bash scripts/gpu_ncu.sh ncu_acq "scripts/acq_step.py 256 1" "acq_score|select_l0|pick_ranks_fast|select_rest|pick_bucket0" 5 > /dev/null 2>&1
# same-box A/B against the round's earlier build (if it travelled) and the other level-0 / scoring kernels
PP_AB_L0=1 PP_AB_SCORE=6 timeout 200 python scripts/acq_ab.py $(ls pixelpick_b200/csrc/build/libpp_prev.so 2>/dev/null) > gpurun_out/acq_ab_final.log 2>&1
( time timeout 600 python bench.py --steps 20 --warmup 5 --sweep ) > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
tail -3 gpurun_out/bench_1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_1.json").read().strip().splitlines()[-1])
    c = d["config"]
    print(d["metric"], d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "acq_step_ms", c["acq_step_ms"],
          "frac", c["acq_step_frac_of_hbm"], "roof", d["roofline"]["frac"], "query", c.get("query_rn50_mpix_s"))
except Exception as e:
    print("bench FAILED", e)
PY
