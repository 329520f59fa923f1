#!/bin/bash
# ncu evidence for one round (run under gpurun; outputs into gpurun_out/).  Usage: tools/profile.sh <tag>
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-extra-configs"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD --skip-e2e > gpurun_out/ncu_launches_$TAG.log 2>&1
for K in pyr_fast_kernel fast_cells_kernel distribute_kernel describe_kernel hamming_topk_kernel match_resolve_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out/
