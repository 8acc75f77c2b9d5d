#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_layers.py tests/test_gpu_fused_epilogues.py tests/test_gpu_stage_abi.py -x -q --timeout 300 > $O/r2epi_pytest.log 2>&1
echo "pytest exit $?"; tail -n 3 $O/r2epi_pytest.log | cut -c1-300
timeout 300 python tools/ab_step.py --toggle operators.EPILOGUE_STAGED 2>/dev/null | tail -1 | tee $O/r2epi_ab.json
