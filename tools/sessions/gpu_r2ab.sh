#!/bin/bash
# tests + step time with / without the side-stream overlap
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r2ab}
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > $O/${T}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 3 $O/${T}_pytest.log | cut -c1-300
for mode in on off; do
  if [ $mode = off ]; then export SN_NO_SIDE_STREAM=1; else unset SN_NO_SIDE_STREAM; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-spmm-sweep --no-e2e > $O/${T}_bench_$mode.json 2> $O/${T}_bench_$mode.err
  echo "bench $mode exit $?"
  python - <<P
import json
d=json.load(open("$O/${T}_bench_$mode.json"))
print("$mode", d["ms_per_step"], d["step_mode"], d["final_loss"], d["clocks"])
P
done
