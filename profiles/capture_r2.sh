#!/bin/bash
# Round-2 ncu evidence (one context of 128 lanes, maps grown over 60+ frames).  usage (under gpurun): profiles/capture_r2.sh <tag>
#   1. launch list of two late frames (gpu__time_duration.sum per launch)            -> gpurun_out/<tag>_launches.csv
#   2. `--set full` of the 5-NN map search, the fit, the association and the LM solve -> gpurun_out/<tag>_*.ncu-rep + raw CSV
tag=${1:-r2a}
mkdir -p gpurun_out
B="python bench.py --lanes 128 --groups 1 --steps 12 --warmup 3 --skip-e2e --no-extras --knn-frames 0 --no-cpu-baseline --no-full-schedule --no-single"
# plain run first (must exit 0 without ncu)
timeout 300 $B > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err; echo "plain rc=$?"
# ~186 launches per frame: frames 70-71 start near launch 13000
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 12900 -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_l.log 2>&1; echo "ncu launches rc=$?"
# k_map_knn + k_map_fit: 20 matching launches per frame; frame 60 -> 1200
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_map_knn|k_map_fit' -s 1200 -c 4 -o gpurun_out/${tag}_map -f $B > gpurun_out/${tag}_ncu_a.log 2>&1; echo "ncu map rc=$?"
# k_odo_assoc_fast / k_odo_assoc<..> / k_lm_solve: 40 per frame -> frame 60 at 2400
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_odo_assoc|k_lm_solve' -s 2400 -c 6 -o gpurun_out/${tag}_odo -f $B > gpurun_out/${tag}_ncu_b.log 2>&1; echo "ncu odo rc=$?"
for f in map odo; do
  [ -f gpurun_out/${tag}_$f.ncu-rep ] && ncu -i gpurun_out/${tag}_$f.ncu-rep --page raw --csv > gpurun_out/${tag}_${f}_raw.csv 2>/dev/null
done
ls -la gpurun_out/ | grep $tag
