#!/bin/bash
# small-operator regime: parity tests + A/B of the small-operator kernel (rg6) against the persistent one (rg7)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_spmm.py -x -q --timeout 300 > $O/r2small_pytest.log 2>&1
echo "pytest exit $?"; tail -n 4 $O/r2small_pytest.log | cut -c1-300
: > $O/r2small_spmm.jsonl
for C in 16 32 64 128 256 512; do
  timeout 300 python tools/spmm_bench.py --meshes 1 --distinct 1 --vertices 7000 --features $C --ops D,Dstar --variants rg,rg6,rg7,direct --reps 30 2>/dev/null | grep -v "copy\|floor" >> $O/r2small_spmm.jsonl
done
timeout 300 python tools/spmm_bench.py --meshes 32 --distinct 32 --vertices 500 --features 128 --ops L --variants rg,rg6,rg7,direct --reps 30 2>/dev/null | grep -v "copy\|floor" >> $O/r2small_spmm.jsonl
for m in 2 4 8 16; do
  timeout 300 python tools/spmm_bench.py --meshes $m --distinct 1 --vertices 7000 --features 128 --ops D,Dstar --variants rg6,rg7 --reps 30 2>/dev/null | grep -v "copy\|floor" >> $O/r2small_spmm.jsonl
done
python - <<'P'
import json
for l in open("gpurun_out/r2small_spmm.jsonl"):
    d=json.loads(l)
    print("%-6s %-7s rows %7d C %3d  %6.2f us (best %6.2f)  frac %.3f" % (d["op"], d["variant"], d["rows"], d["C"], d["us"], d["us_best"], d["frac_of_measured_peak"]))
P
timeout 300 python tools/spmm_bench.py --meshes 1 --distinct 1 --vertices 7000 --features 16 --ops D --variants rg --reps 50 2>/dev/null | grep floor
