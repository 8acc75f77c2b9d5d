#!/bin/bash
# ncu --set full of the small glue kernels inside one eager training step (first launch of each)
set -u
mkdir -p gpurun_out
O=gpurun_out
for k in avg_fold_fwd_kernel avg_fold_bwd_kernel bn_fold_fwd_kernel bn_fold_bwd_kernel segment_sum_kernel avg_stats_kernel colstats_final_kernel reduce_partials4_kernel avg_pre_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -f --profile-from-start off -k regex:$k -s 2 -c 1 -o $O/r2tiny_$k python tools/step_once.py > $O/r2tiny_$k.log 2>&1
  echo "$k $?"
done
ls -la $O/r2tiny_*.ncu-rep
