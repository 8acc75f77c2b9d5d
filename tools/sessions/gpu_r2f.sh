#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=r2f
timeout 900 python -m pytest tests/test_gpu_fused_epilogues.py tests/test_gpu_spmm.py -x -q --timeout 300 > $O/${T}_pytest_fused.log 2>&1
echo "pytest fused exit $?"; tail -n 5 $O/${T}_pytest_fused.log | cut -c1-400
timeout 600 python tools/fusion_bench.py > $O/${T}_fusion.json 2> $O/${T}_fusion.err
echo "fusion exit $?"; python -c "
import json
for k,v in json.load(open('$O/${T}_fusion.json')).items(): print('%-50s %s'%(k,v))"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-spmm-sweep > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
echo "bench exit $?"; cut -c1-200 $O/${T}_bench_n1.json
