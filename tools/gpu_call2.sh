#!/bin/bash
# Round-1 GPU session A: row-group SpMM kernel -- operator sweep, parity tests, training-step bench, one ncu capture.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/b_smi.txt 2>&1
echo "== spmm cfg3 C=128" > $O/b_spmm.log
timeout 240 python tools/spmm_bench.py --reps 20 >> $O/b_spmm.log 2>&1
echo "== spmm cfg3 C=256" >> $O/b_spmm.log
timeout 180 python tools/spmm_bench.py --reps 10 --features 256 --ops D,Dstar --variants rg,rg1,rg2,rg3,rg4,smem >> $O/b_spmm.log 2>&1
echo "== spmm cfg2 (32 x 500 V) C=128" >> $O/b_spmm.log
timeout 120 python tools/spmm_bench.py --reps 20 --meshes 32 --vertices 500 --variants rg,rg1,rg2,rg3,smem,direct >> $O/b_spmm.log 2>&1
echo "== spmm cfg5 (1 x 7000 V) C=256" >> $O/b_spmm.log
timeout 120 python tools/spmm_bench.py --reps 20 --meshes 1 --distinct 1 --vertices 7000 --features 256 --variants rg,rg1,rg2,rg3,smem,direct >> $O/b_spmm.log 2>&1
echo "== spmm cfg3 C=64 / C=16" >> $O/b_spmm.log
timeout 120 python tools/spmm_bench.py --reps 10 --features 64 --variants rg,rg1,rg2,rg3,direct >> $O/b_spmm.log 2>&1
timeout 120 python tools/spmm_bench.py --reps 10 --features 16 --variants rg,rg1,rg2,rg3,direct >> $O/b_spmm.log 2>&1
tail -n 60 $O/b_spmm.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/b_pytest.log 2>&1
echo "pytest exit $?" >> $O/b_pytest.log
tail -n 15 $O/b_pytest.log
timeout 420 python bench.py --steps 10 --warmup 3 > $O/b_bench_n1.json 2> $O/b_bench_n1.err
echo "bench exit $?"
cut -c1-1500 $O/b_bench_n1.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowgroup -c 10 -f -o $O/b_rg_full \
  python tools/spmm_bench.py --reps 1 --ops D,Dstar --variants rg > $O/b_ncu.log 2>&1
echo "ncu exit $?"
ls -la $O
