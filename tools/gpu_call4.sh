#!/bin/bash
# Round-1 GPU session D: final row-group defaults (bench + tests) and a look at the tcgen05 GEMM (timings + ncu).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/gemm_bench.py > $O/d_gemm.log 2>&1
cat $O/d_gemm.log | cut -c1-700
echo "== spmm cfg3 C=128" > $O/d_spmm.log
timeout 240 python tools/spmm_bench.py --reps 20 --variants rg,rg1,rg2,rg3,rg4,rg5,smem >> $O/d_spmm.log 2>&1
echo "== spmm cfg3 C=512 / 64" >> $O/d_spmm.log
timeout 120 python tools/spmm_bench.py --reps 10 --features 512 --meshes 32 --variants rg,rg1,rg2,rg3 >> $O/d_spmm.log 2>&1
timeout 120 python tools/spmm_bench.py --reps 10 --features 64 --variants rg,rg1,rg2,rg3 >> $O/d_spmm.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/d_pytest.log 2>&1
echo "pytest exit $?" >> $O/d_pytest.log
tail -n 4 $O/d_pytest.log
timeout 420 python bench.py --steps 10 --warmup 3 > $O/d_bench_n1.json 2> $O/d_bench_n1.err
echo "bench exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -c 8 -f -o $O/d_gemm_full \
  python tools/gemm_bench.py > $O/d_ncu.log 2>&1
echo "ncu exit $?"
