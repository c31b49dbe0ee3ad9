#!/bin/bash
# A/B timing aid: bench.py (device-resident + e2e arms only) over library builds / context splits.
# usage: ab_run.sh "<lib or ->:<lanes>:<groups>[:<graphs>]" ...   -> one short line per run in gpurun_out/ab.log
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read lib lanes groups graphs <<< "$spec"
  if [ "$lib" != "-" ]; then export LVO_LIB_PATH=$PWD/lidar-visual-odometry_b200/ab/$lib; else unset LVO_LIB_PATH; fi
  python bench.py --no-extras --knn-frames 0 --no-cpu-baseline --steps 10 --warmup 3 --lanes $lanes --groups $groups --graphs ${graphs:--1} $LVO_AB_EXTRA 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=lambda x: x if x!=x else round(x); print('$spec', 'value', r(d['value']), 'e2e', r(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'knn_us', round(d['roofline']['avg_launch_us'],1))" | tee -a gpurun_out/ab.log
done
