#!/bin/bash
# ncu evidence of the shipped kernels at FULL batch (256 frames / 2048 pairs per launch); outputs into gpurun_out/.
# Usage (under gpurun): tools/profile_r2.sh <tag>
TAG=${1:-r2}
T="python tools/full_batch_pass.py"
$T match > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $T match > gpurun_out/ncu_launches_$TAG.log 2>&1
cap() {  # kernel regex, launches to skip, launches to capture, extra args of the target
  ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/prof_$1_$TAG $T $4 > gpurun_out/ncu_$1_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_$1_$TAG.log
}
cap "pyr_fast_kernel" 24 8 ""          # one whole step: blur of level 0 + 7 fused resize + blur launches
cap fast_cells_kernel 3 1 ""
cap distribute_kernel 3 1 ""
cap describe_kernel 3 1 ""
cap hamming_topk_kernel 3 1 match
cap match_resolve_kernel 3 1 match
ls -la gpurun_out/*_$TAG.ncu-rep
