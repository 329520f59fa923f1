#!/bin/bash
# ncu --set full capture of one kernel (regex $1) of the bench; tag $2; skip $3 launches.
K=$1; TAG=${2:-x}; SKIP=${3:-8}
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
tail -3 gpurun_out/ncu_${K}_$TAG.log
