#!/bin/bash
# gpurun --gpus N -- 'bash scripts/gpu_dp_ab.sh N'  : the DP train leg under several reducer settings
mkdir -p gpurun_out
N=$1
run() {
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29545 \
      bench.py --gpus "$N" --no-extras --no-query --no-cpu-baseline > gpurun_out/dp_${N}_$name.json 2> gpurun_out/dp_${N}_$name.err
  python - "$N" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/dp_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(sys.argv[2], "N", sys.argv[1], "img/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run default PP_X=0
run no_overlap PP_DP_OVERLAP=0
run nocomm PP_DP_NOCOMM=1
run bucket8 PP_DP_BUCKET_MB=8
run bucket128 PP_DP_BUCKET_MB=128
