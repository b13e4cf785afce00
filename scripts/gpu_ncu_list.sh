#!/bin/bash
# gpurun -- 'bash scripts/gpu_ncu_list.sh NAME "<script + args>" [launch count]'  -> per-launch durations (csv) + a full capture
mkdir -p gpurun_out
NAME=$1; CMD=$2; C=${3:-400}
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c $C --csv --log-file gpurun_out/${NAME}_launches.csv python $CMD > gpurun_out/${NAME}_run.log 2>&1
tail -2 gpurun_out/${NAME}_run.log
python - "$NAME" <<'PY'
import csv, sys, collections
name = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/{name}_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[ii], r[ki][:70]), {})[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.OrderedDict()
for (i, k), m in per.items():
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += m.get("gpu__time_duration.sum", 0); a[2] += m.get("dram__bytes_read.sum", 0); a[3] += m.get("dram__bytes_write.sum", 0)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} n={a[0]:4d} avg {a[1] / a[0] / 1e3:9.2f} us  dram rd {a[2] / a[0] / 1e6:9.2f} MB wr {a[3] / a[0] / 1e6:9.2f} MB")
PY
