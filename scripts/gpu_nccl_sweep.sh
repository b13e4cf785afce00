#!/bin/bash
# gpurun --gpus N -- 'bash scripts/gpu_nccl_sweep.sh N "<bench flags>" v1 v2 ...'  : train leg at N GPUs for several NCCL_MAX_CTAS values
mkdir -p gpurun_out
N=$1; FLAGS=$2; shift; shift
for v in "$@"; do
  if [ "$v" = "default" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$v; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29544 \
      bench.py --gpus "$N" $FLAGS > gpurun_out/sweep_${N}_$v.json 2> gpurun_out/sweep_${N}_$v.err
  python - "$N" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/sweep_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print("NCCL_MAX_CTAS", sys.argv[2], "N", sys.argv[1], "img/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
