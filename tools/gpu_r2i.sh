#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r2i}
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > $O/${T}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $O/${T}_pytest.log
tail -n 12 $O/${T}_pytest.log | cut -c1-400
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-spmm-sweep > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
echo "bench exit $?"; cut -c1-200 $O/${T}_bench_n1.json; tail -n 3 $O/${T}_bench_n1.err | cut -c1-300
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_raw.csv python tools/step_once.py > $O/${T}_step_once.log 2>&1
echo "ncu exit $?"; tail -n 1 $O/${T}_step_once.log | cut -c1-200
python tools/launch_list.py $O/${T}_launches_raw.csv $O/${T}_launches_step | head -40
