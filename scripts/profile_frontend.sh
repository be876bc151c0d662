#!/bin/bash
# ncu captures of the front-end kernel (run under gpurun).  $1 = tag, $2 = kernel regex
TAG=${1:-r01}
KRN=${2:-frontend}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$KRN -s 3 -c 1 \
    -o gpurun_out/prof_frontend_$TAG -f python scripts/quick_time.py 256 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
