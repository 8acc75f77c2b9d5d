#!/bin/bash
# batch-assembly plan: assembly parity test + the bench's end-to-end loops
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_spmm.py -x -q --timeout 200 -k "assembly" > $O/r2e2e_pytest.log 2>&1
echo "pytest exit $?"; tail -n 2 $O/r2e2e_pytest.log | cut -c1-200
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-spmm-sweep > $O/r2e2e_bench_n1.json 2> $O/r2e2e_bench_n1.err
echo "bench exit $?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r2e2e_bench_n1.json"))
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "coo", d["e2e_coo_upload"]["ms_per_step"], "built", d["e2e_gpu_built_operators"]["ms_per_step"])
P
