#!/bin/bash
# ncu evidence of one version: launch list (time per launch) + `--set full` of the hot kernels, 128 lanes in one context.
# usage (under gpurun): profiles/capture_r1.sh <tag>
tag=${1:-s5}
B="python bench.py --lanes 128 --groups 1 --steps 2 --warmup 3 --skip-e2e --no-extras --knn-frames 0 --no-cpu-baseline"
$B > gpurun_out/plain_$tag.json 2> gpurun_out/plain_$tag.log || exit 1     # un-profiled first: must exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches_${tag}_l128.csv $B > gpurun_out/ncu_list_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_sector_sort|k_lessflat_voxel|k_vx_centroid|k_map_knn|k_map_fit' -s 96 -c 6 \
  -o gpurun_out/prof_${tag}_extract_map -f $B > gpurun_out/ncu_a_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_odo_assoc|k_lm_solve' -s 163 -c 4 \
  -o gpurun_out/prof_${tag}_odo_lm -f $B > gpurun_out/ncu_b_$tag.log 2>&1
ls -la gpurun_out/*${tag}*
