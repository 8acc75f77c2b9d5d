#!/bin/bash
# Round-1 GPU session C: row-group SpMM pipeline-shape variants (stages in flight x entries per stage x CTAs per SM).
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== spmm cfg3 C=128" > $O/c_spmm.log
timeout 240 python tools/spmm_bench.py --reps 20 --variants rg,rg3,rg4,rg5,rg6,rg7,rg8,rg9,smem >> $O/c_spmm.log 2>&1
echo "== spmm cfg3 C=256" >> $O/c_spmm.log
timeout 180 python tools/spmm_bench.py --reps 10 --features 256 --variants rg,rg4,rg5,rg6,rg7,smem >> $O/c_spmm.log 2>&1
echo "== spmm cfg3 C=64 / 32 / 512" >> $O/c_spmm.log
timeout 120 python tools/spmm_bench.py --reps 10 --features 64 --variants rg,rg1,rg2,rg3,direct >> $O/c_spmm.log 2>&1
timeout 120 python tools/spmm_bench.py --reps 10 --features 32 --variants rg,rg1,rg2,rg3,direct >> $O/c_spmm.log 2>&1
timeout 120 python tools/spmm_bench.py --reps 10 --features 512 --meshes 32 --variants rg,rg2,rg3,smem,direct >> $O/c_spmm.log 2>&1
echo "== spmm cfg2 (32 x 500 V) C=128" >> $O/c_spmm.log
timeout 120 python tools/spmm_bench.py --reps 20 --meshes 32 --vertices 500 --variants rg,rg1,rg2,rg5,rg6,direct >> $O/c_spmm.log 2>&1
grep -c variant $O/c_spmm.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/c_pytest.log 2>&1
echo "pytest exit $?" >> $O/c_pytest.log
tail -n 6 $O/c_pytest.log
timeout 420 python bench.py --steps 10 --warmup 3 > $O/c_bench_n1.json 2> $O/c_bench_n1.err
echo "bench exit $?"
cut -c1-300 $O/c_bench_n1.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowgroup -c 10 -f -o $O/c_rg_full \
  python tools/spmm_bench.py --reps 1 --ops D,Dstar --variants rg > $O/c_ncu.log 2>&1
echo "ncu exit $?"
