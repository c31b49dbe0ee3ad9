#!/bin/bash
# Launch list only (gpu__time_duration.sum per launch) of the two isolated full-schedule frames after a short 128-lane run.
# usage (under gpurun): profiles/launchlist_r2.sh <tag>
tag=${1:-r2g}
mkdir -p gpurun_out
B="python bench.py --lanes 128 --groups 1 --steps 4 --warmup 3 --skip-e2e --no-extras --knn-frames 0 --no-cpu-baseline --no-full-schedule --no-single"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $B --profile-iso 2 > gpurun_out/${tag}_ncu_l.log 2>&1; echo "ncu launches rc=$?"
python profiles/launch_shares.py gpurun_out/${tag}_launches.csv 1 16
