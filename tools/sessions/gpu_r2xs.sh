#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmm.py -x -q --timeout 300 -k "rowgroup or widths" > $O/r2xs_pytest.log 2>&1
echo "pytest exit $?"; tail -n 3 $O/r2xs_pytest.log | cut -c1-300
timeout 300 python tools/spmm_bench.py --ops D,Dstar,L --variants rg,rg9,rg10,rg5 --reps 30 2>/dev/null | grep -v "copy\|floor" > $O/r2xs_spmm.jsonl
timeout 300 python tools/spmm_bench.py --ops D,Dstar --features 256 --variants rg,rg9,rg10 --reps 30 2>/dev/null | grep -v "copy\|floor" >> $O/r2xs_spmm.jsonl
python - <<'P'
import json
for l in open("gpurun_out/r2xs_spmm.jsonl"):
    d=json.loads(l)
    print("%-6s %-7s rows %7d C %3d  %6.2f us (best %6.2f)  frac %.3f" % (d["op"], d["variant"], d["rows"], d["C"], d["us"], d["us_best"], d["frac_of_measured_peak"]))
P
timeout 300 python tools/ab_step.py --toggle operators.STATS_XS 2>/dev/null | tail -1 | tee $O/r2xs_ab.json
