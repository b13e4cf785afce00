#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_ncu.sh NAME "<python script + args>" [kernel regex] [launch count]'
mkdir -p gpurun_out
NAME=$1; CMD=$2; K=${3:-"conv_igemm|wgrad"}; C=${4:-12}
timeout 800 ncu --set full --clock-control none --import-source on -k "regex:$K" -c $C -o gpurun_out/$NAME -f python $CMD > gpurun_out/${NAME}_run.log 2>&1
tail -3 gpurun_out/${NAME}_run.log
ncu -i gpurun_out/$NAME.ncu-rep --page details --csv > gpurun_out/${NAME}_details.csv 2>/dev/null
ls -la gpurun_out/$NAME.ncu-rep gpurun_out/${NAME}_details.csv
