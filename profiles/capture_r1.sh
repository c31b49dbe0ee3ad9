#!/bin/bash
# ncu evidence of one version: `--set full` of the hot kernels of a steady-state frame (128 lanes in one context), raw pages as CSV.
# usage (under gpurun): profiles/capture_r1.sh <tag>      (the launch list comes from profiles/final_r1.sh)
tag=${1:-s6}
mkdir -p gpurun_out
timeout 60 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
B="python bench.py --lanes 128 --groups 1 --steps 2 --warmup 3 --skip-e2e --no-extras --knn-frames 0 --no-cpu-baseline --no-full-schedule"
# matching launches per frame: sector_sort 1, lessflat_voxel 1, vx_centroid 2, map_knn 10, map_fit 10 -> frame 4 starts at 96
timeout 100 ncu --set full --clock-control none --import-source on -k 'regex:k_sector_sort|k_lessflat_voxel|k_vx_centroid|k_map_knn|k_map_fit' -s 96 -c 6 \
  -o gpurun_out/prof_${tag}_extract_map -f $B > gpurun_out/ncu_a_$tag.log 2>&1; echo "ncu a rc=$?"
# assoc_fast 10, assoc<8|32> 10, lm_solve 20 per frame -> frame 4 starts at 160; 163: lm(o=0, map) is skipped, first capture is in outer iteration 1
timeout 100 ncu --set full --clock-control none --import-source on -k 'regex:k_odo_assoc|k_lm_solve' -s 161 -c 5 \
  -o gpurun_out/prof_${tag}_odo_lm -f $B > gpurun_out/ncu_b_$tag.log 2>&1; echo "ncu b rc=$?"
for f in extract_map odo_lm; do
  [ -f gpurun_out/prof_${tag}_$f.ncu-rep ] && ncu -i gpurun_out/prof_${tag}_$f.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_${f}_raw.csv 2>/dev/null
done
ls -la gpurun_out/ | grep $tag
