#!/bin/bash
# round 2, session d: CapturedTrainStep in the package, bench refactor, smoke at C=128
set -u
mkdir -p gpurun_out
O=gpurun_out
T=r2d
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${T}_smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1
echo "smoke exit $?"; tail -n 2 $O/${T}_smoke.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > $O/${T}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $O/${T}_pytest.log
tail -n 15 $O/${T}_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
echo "bench exit $?"; cut -c1-600 $O/${T}_bench_n1.json; tail -n 5 $O/${T}_bench_n1.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
echo "ref exit $?"; cut -c1-1200 $O/${T}_bench_ref.json; tail -n 3 $O/${T}_bench_ref.err | cut -c1-300
