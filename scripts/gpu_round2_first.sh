#!/bin/bash
# First gpurun call of the next round: confirms everything written at the end of round 1 without GPU time
# (sharded QuerySelector / torchrun main_al, AcqSession buffer checks, index clamps, --overlap-select) and refreshes the bench.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first.sh'            (1 GPU)
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_round2_first.sh 2'  (the torchrun legs)
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  ( time timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --durations=8 ) > gpurun_out/pytest_full.log 2>&1
  tail -14 gpurun_out/pytest_full.log > gpurun_out/pytest.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
  ( time timeout 900 python bench.py --no-train --no-cpu-baseline ) > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
  ( time timeout 900 python bench.py --no-train --no-cpu-baseline --overlap-select ) > gpurun_out/bench_q_overlap.json 2> gpurun_out/bench_q_overlap.err
  # the active-learning loop end to end on synthetic data (graph path)
  ( time timeout 600 python -m pixelpick_b200.main_al --dataset_name cs --dir_root gpurun_out/al1 --n_workers 0 --synthetic 16 256 512 \
      --n_epochs 2 --max_budget 20 ) > gpurun_out/main_al_1gpu.log 2>&1
  tail -4 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log
  for f in bench_q bench_q_overlap; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 4), "score frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
  tail -3 gpurun_out/main_al_1gpu.log
else
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
      -m pixelpick_b200.main_al --dataset_name cs --dir_root gpurun_out/al$N --n_workers 0 --synthetic 16 256 512 --n_epochs 2 \
      --max_budget 20 ) > gpurun_out/main_al_${N}gpu.log 2>&1
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus "$N" --no-train ) > gpurun_out/bench_${N}gpu_q.json 2> gpurun_out/bench_${N}gpu_q.err
  tail -5 gpurun_out/main_al_${N}gpu.log; head -c 400 gpurun_out/bench_${N}gpu_q.json
  # the per-rank pick files must equal a single-process run: compare queries.pkl of the two runs if both exist
  python - "$N" <<'PY'
import glob, pickle, sys
import numpy as np
a = sorted(glob.glob("gpurun_out/al1/checkpoints/*/0_query/queries.pkl"))
b = sorted(glob.glob(f"gpurun_out/al{sys.argv[1]}/checkpoints/*/0_query/queries.pkl"))
print("queries.pkl files:", a, b)
PY
fi
