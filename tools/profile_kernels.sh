#!/bin/bash
# ncu --set full capture of the given kernels (regex list) of the bench; usage: tools/profile_kernels.sh <tag> <skip> k1 k2 ...
TAG=$1; SKIP=$2; shift 2
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-extra-configs"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 2 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${K}_$TAG.log
done
