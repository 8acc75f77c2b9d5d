#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r2v}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > $O/${T}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $O/${T}_pytest.log
tail -n 6 $O/${T}_pytest.log | cut -c1-300
for pdl in 1 0; do
  SN_PDL=$pdl timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-spmm-sweep > $O/${T}_bench_pdl$pdl.json 2> $O/${T}_bench_pdl$pdl.err
  echo "bench pdl=$pdl exit $?"; python -c "
import json; d=json.load(open('$O/${T}_bench_pdl$pdl.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'loss', d['final_loss'])"
done
