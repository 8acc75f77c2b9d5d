#!/bin/bash
# Round-1 GPU session E: GPU operator construction (f3), seam mirror, Siamese; final row-group ncu capture; launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/e_pytest.log 2>&1
echo "pytest exit $?" >> $O/e_pytest.log
tail -n 25 $O/e_pytest.log | cut -c1-250
timeout 300 python tools/mesh_ops_bench.py > $O/e_mesh_ops.log 2>&1
tail -n 2 $O/e_mesh_ops.log | cut -c1-600
timeout 200 python tools/spmm_bench.py --reps 10 --features 512 --meshes 32 --variants rg,rg1,rg2 > $O/e_spmm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowgroup -c 10 -f -o $O/e_rg_full \
  python tools/spmm_bench.py --reps 1 --ops D,Dstar --variants rg > $O/e_ncu.log 2>&1
echo "ncu exit $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/e_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-spmm-sweep > $O/e_launch_bench.log 2>&1
echo "launch list exit $?"; wc -l $O/e_launches.csv
