#!/bin/bash
# Launch list (gpu__time_duration.sum per launch) of late frames of a short 128-lane run.  ~190 launches per frame; 24 frames in all
# (3 warm-up + 1 timed step of 5 frames each, then 4 isolated frames with the fixed-point skip off = the full ten-iteration schedule).
# usage (under gpurun): profiles/launchlist_r2.sh <tag>
tag=${1:-r2b}
mkdir -p gpurun_out
B="python bench.py --lanes 128 --groups 1 --steps 1 --warmup 3 --skip-e2e --no-extras --knn-frames 0 --no-cpu-baseline --no-full-schedule --no-single"
timeout 200 $B > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err; echo "plain rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 3200 -c 1400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_l.log 2>&1; echo "ncu launches rc=$?"
python profiles/launch_shares.py gpurun_out/${tag}_launches.csv 2 30
