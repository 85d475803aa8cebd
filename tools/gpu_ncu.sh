#!/bin/bash
# ncu session: launch list + one full capture of the hot kernel.  $1 = extra args of tools/ncu_run.py, $2 = tag
mkdir -p gpurun_out
TAG=${2:-ncu}
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/ncu_run.py --reps 3 $1 > gpurun_out/${TAG}_launches.log 2>&1; echo "launches rc=$?"; tail -2 gpurun_out/${TAG}_launches.log
ncu --set full --clock-control none --import-source on -k regex:fast_cost -s 1 -c 1 -o gpurun_out/${TAG}_prof -f python tools/ncu_run.py --reps 2 $1 > gpurun_out/${TAG}_prof.log 2>&1; echo "full rc=$?"; tail -2 gpurun_out/${TAG}_prof.log
ls -la gpurun_out/
