#!/bin/bash
# Round-1 GPU session H: elu' folded into the dZ GEMM epilogue and the backward SpMM store path; tests + bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/h_pytest.log 2>&1
echo "pytest exit $?" >> $O/h_pytest.log
tail -n 30 $O/h_pytest.log | cut -c1-250
timeout 420 python bench.py --steps 10 --warmup 3 > $O/h_bench_n1.json 2> $O/h_bench_n1.err
echo "bench n1 exit $?"
cut -c1-200 $O/h_bench_n1.json
tail -n 3 $O/h_bench_n1.err | cut -c1-300
