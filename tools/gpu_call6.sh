#!/bin/bash
# Round-1 GPU session F: L2 prefetch in the tcgen05 GEMMs (A/B), colstats final reduce, full tests + bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tools/gemm_bench.py > $O/f_gemm.log 2>&1
grep -c "^{" $O/f_gemm.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/f_pytest.log 2>&1
echo "pytest exit $?" >> $O/f_pytest.log
tail -n 4 $O/f_pytest.log | cut -c1-200
timeout 420 python bench.py --steps 10 --warmup 3 > $O/f_bench_n1.json 2> $O/f_bench_n1.err
echo "bench exit $?"
cut -c1-200 $O/f_bench_n1.json
