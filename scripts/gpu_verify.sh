#!/bin/bash
# full GPU suite + smoke + the default bench line (what the driver runs at round end)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=8 ) > gpurun_out/pytest_full.log 2>&1
tail -14 gpurun_out/pytest_full.log > gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 $BENCH_ARGS ) > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
timeout 300 python scripts/profile_step.py resnet 32 > gpurun_out/prof_step_rn32.log 2>&1
tail -4 gpurun_out/pytest.log; tail -1 gpurun_out/smoke.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_1.json").read().strip().splitlines()[-1])
print(d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
print({k: v for k, v in d["config"].items() if not isinstance(v, (dict, list))})
PY
