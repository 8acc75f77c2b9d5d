#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=r2h
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_raw.csv python tools/step_once.py > $O/${T}_step_once.log 2>&1
echo "ncu exit $?"; tail -n 2 $O/${T}_step_once.log | cut -c1-200
python tools/launch_list.py $O/${T}_launches_raw.csv $O/${T}_launches_step
