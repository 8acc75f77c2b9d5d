#!/bin/bash
# A/B of the mesh renumbering (geometry.locality_order): operator-only times and the whole step
set -u
mkdir -p gpurun_out
O=gpurun_out
for ord in none bisect morton_xy rcm; do
  timeout 300 python tools/spmm_bench.py --ops D,Dstar --variants rg --order $ord --distinct 16 2>/dev/null | grep -v copy | cut -c1-230
done > $O/r2order_spmm.jsonl
cat $O/r2order_spmm.jsonl
for ord in none bisect; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-spmm-sweep --no-e2e --mesh-order $ord > $O/r2order_bench_$ord.json 2> $O/r2order_bench_$ord.err
  echo "bench $ord exit $?"
  python - <<P
import json
d=json.load(open("$O/r2order_bench_$ord.json"))
print("$ord", d["ms_per_step"], d["roofline_spmm"]["frac"], d["roofline_spmm"]["avg_launch_us"])
for k,v in d["kernels"].items():
    if "spmm" in k: print("   ",k, {a:b for a,b in v.items() if a in ("launches","avg_us","ms_per_step","GBps")})
P
done
