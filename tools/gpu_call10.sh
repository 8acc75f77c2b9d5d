#!/bin/bash
# Round-1 GPU session J: single-node Dirac block (gradient accumulation in epilogues).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/j_pytest.log 2>&1
echo "pytest exit $?" >> $O/j_pytest.log
tail -n 30 $O/j_pytest.log | cut -c1-250
timeout 420 python bench.py --steps 10 --warmup 3 > $O/j_bench_n1.json 2> $O/j_bench_n1.err
echo "bench n1 exit $?"
cut -c1-200 $O/j_bench_n1.json
tail -n 3 $O/j_bench_n1.err | cut -c1-300
