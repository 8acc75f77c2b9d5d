#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=r2q
timeout 900 python -m pytest tests/test_gpu_gemm.py -x -q --timeout 300 -k "tn" > $O/${T}_pytest_tn.log 2>&1
echo "pytest tn exit $?"; tail -n 5 $O/${T}_pytest_tn.log | cut -c1-300
timeout 600 python tools/gemm_bench.py > $O/${T}_gemm.log 2>&1
grep gemm_tn $O/${T}_gemm.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:(round(v['us'],1) if isinstance(v,dict) else v) for k,v in d.items()})"
