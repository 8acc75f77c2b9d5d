#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=r2l
timeout 600 python tools/gemm_stage_bench.py > $O/${T}_gemm_stage.log 2>&1
echo "exit $?"; cat $O/${T}_gemm_stage.log | cut -c1-600
timeout 300 python -m pytest tests/test_gpu_head_tail.py tests/test_gpu_gemm.py -x -q --timeout 300 2>&1 | tail -n 3
