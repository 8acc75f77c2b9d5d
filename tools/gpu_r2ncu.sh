#!/bin/bash
# ncu --set full captures of the round-2 kernels (one launch each), for profiles/r2_*_ncu_summary.json
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:gemm_tf32_ts_kernel -s 1 -c 1 -o $O/r2_gemm_ts python tools/gemm_bench.py --ncu > $O/r2_ncu_gemm.log 2>&1; echo "gemm $?"
timeout 600 $NCU -k regex:gemm_tn_ts_kernel -s 1 -c 1 -o $O/r2_gemm_tn python tools/gemm_bench.py --ncu > $O/r2_ncu_tn.log 2>&1; echo "tn $?"
timeout 600 $NCU -k regex:gemm_tf32_ts_kernel -s 1 -c 1 -o $O/r2_gemm_act python tools/fusion_bench.py --ncu-act > $O/r2_ncu_act.log 2>&1; echo "act $?"
timeout 600 $NCU -k regex:rowgroup_spmm_kernel -s 4 -c 4 -o $O/r2_rowgroup_block python tools/block_step.py > $O/r2_ncu_block.log 2>&1; echo "block $?"
ls -la $O/r2_*.ncu-rep
