#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=r2g
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_ts_kernel -s 1 -c 1 -f -o $O/${T}_gemm_act python tools/fusion_bench.py --ncu-act > $O/${T}_ncu_act.log 2>&1
echo "ncu act exit $?"; tail -n 2 $O/${T}_ncu_act.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowgroup_spmm_kernel -s 1 -c 1 -f -o $O/${T}_spmm_stats python tools/fusion_bench.py --ncu-spmm-stats > $O/${T}_ncu_spmm.log 2>&1
echo "ncu spmm exit $?"; tail -n 2 $O/${T}_ncu_spmm.log | cut -c1-200
