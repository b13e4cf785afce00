#!/bin/bash
# gpurun [--gpus N] --timeout 1500 -- 'bash scripts/gpu_bench.sh N [extra bench flags]'  -> gpurun_out/bench_N.json (+ .err)
mkdir -p gpurun_out
N=${1:-1}; shift
if [ "$N" = "1" ]; then
  ( time timeout 1200 python bench.py --gpus 1 "$@" ) > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
else
  ( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus "$N" "$@" ) > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
fi
tail -c 1800 gpurun_out/bench_$N.json; echo; tail -8 gpurun_out/bench_$N.err
