#!/bin/bash
# gpurun --gpus 2: the NCCL data-parallel equality test, then the bench line at N = 2 (as the driver launches it)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -s --timeout 500 ) > gpurun_out/test_dp.log 2>&1
grep -n "passed\|failed\|error\|skipped" gpurun_out/test_dp.log | tail -3
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras ) > gpurun_out/bench_2.json 2> gpurun_out/bench_2.err
tail -c 1800 gpurun_out/bench_2.json; echo; tail -3 gpurun_out/bench_2.err
