#!/bin/bash
# Round-1 GPU session N: AvgResNet2 residual gradient in the ELU-backward kernel, persistent buffers for per-step operator construction.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/n_smoke.log 2>&1
tail -n 2 $O/n_smoke.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/n_pytest.log 2>&1
echo "pytest exit $?" >> $O/n_pytest.log
tail -n 3 $O/n_pytest.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 > $O/n_bench_n1.json 2> $O/n_bench_n1.err
echo "bench n1 exit $?"
cut -c1-160 $O/n_bench_n1.json
tail -n 2 $O/n_bench_n1.err | cut -c1-300
timeout 120 python tools/spmn_bench.py --reps 20 --meshes 1 --distinct 1 --vertices 7000 --features 256 --variants rg,rg1,direct > $O/k_spmm.log 2>&1
timeout 120 python tools/spmn_bench.py --reps 20 --meshes 1 --distinct 1 --vertices 7000 --features 512 --variants rg,rg1,direct >> $O/k_spmm.log 2>&1
