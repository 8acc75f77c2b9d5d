#!/bin/bash
# Round-end style session on one B200: smoke(), the GPU parity tests, the full default bench line, the reference arm, and the
# ncu launch list of one eager step (profiles/r2_launches_step*).
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r2final}
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1
tail -n 1 $O/${T}_smoke.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/${T}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 3 $O/${T}_pytest.log | cut -c1-200
timeout 900 python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
echo "bench exit $?"; cut -c1-250 $O/${T}_bench_n1.json; tail -n 2 $O/${T}_bench_n1.err | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
echo "reference exit $?"; cut -c1-300 $O/${T}_bench_reference.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_raw.csv python tools/step_once.py > $O/${T}_step_once.log 2>&1
echo "ncu exit $?"
python tools/launch_list.py $O/${T}_launches_raw.csv $O/${T}_launches_step | head -32
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:rowgroup_spmm_kernel -s 4 -c 4 -o $O/${T}_rowgroup_block python tools/block_step.py > $O/${T}_ncu_block.log 2>&1
echo "ncu block $?"
