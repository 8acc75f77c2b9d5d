#!/bin/bash
# round 2, session e: fused epilogues (GEMM activation + statistics, SpMM statistics), chained Dirac blocks
set -u
mkdir -p gpurun_out
O=gpurun_out
T=r2e
timeout 900 python -m pytest tests/test_gpu_fused_epilogues.py -x -q --timeout 300 > $O/${T}_pytest_fused.log 2>&1
echo "pytest fused exit $?"; tail -n 25 $O/${T}_pytest_fused.log | cut -c1-400
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > $O/${T}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $O/${T}_pytest.log
tail -n 15 $O/${T}_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
echo "bench exit $?"; cut -c1-300 $O/${T}_bench_n1.json; tail -n 5 $O/${T}_bench_n1.err | cut -c1-300
