#!/bin/bash
# Round-1 GPU session I: epilogue operand prefetch, transposes from the mesh kernels, e2e from geometry.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/i_pytest.log 2>&1
echo "pytest exit $?" >> $O/i_pytest.log
tail -n 30 $O/i_pytest.log | cut -c1-250
timeout 420 python bench.py --steps 10 --warmup 3 > $O/i_bench_n1.json 2> $O/i_bench_n1.err
echo "bench n1 exit $?"
cut -c1-200 $O/i_bench_n1.json
tail -n 3 $O/i_bench_n1.err | cut -c1-300
