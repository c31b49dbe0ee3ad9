#!/bin/bash
# Round-2 ncu evidence.  One context of 128 lanes runs 35 frames unprofiled (the maps grow), then bench.py brackets the isolated
# full-schedule frames (LVO_OPT_FIXPOINT_SKIP = 0: all ten outer iterations for every lane) with cudaProfilerStart / Stop, so that
# `ncu --profile-from-start off` neither profiles nor serialises anything before them.
#   1. launch list of two frames (gpu__time_duration.sum per launch)   -> gpurun_out/<tag>_launches.csv
#   2. `--set full` of the ten 5-NN launches of one frame               -> gpurun_out/<tag>_knn.ncu-rep  + raw CSV
#   3. `--set full` of the other hot kernels of one frame               -> gpurun_out/<tag>_hot.ncu-rep  + raw CSV
# usage (under gpurun): profiles/capture_r2.sh <tag>
tag=${1:-r2h}
mkdir -p gpurun_out
B="python bench.py --lanes 128 --groups 1 --steps 4 --warmup 3 --skip-e2e --no-extras --knn-frames 0 --no-cpu-baseline --no-full-schedule --no-single"
timeout 300 $B > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err; echo "plain rc=$?"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $B --profile-iso 2 > gpurun_out/${tag}_ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:k_map_knn' -c 10 -o gpurun_out/${tag}_knn -f $B --profile-iso 1 > gpurun_out/${tag}_ncu_a.log 2>&1; echo "ncu knn rc=$?"
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:k_odo_assoc|k_lm_solve|k_lessflat_voxel|k_sector_sort|k_sort_onesweep|k_map_fit|k_classify|k_ring_scatter' -c 40 -o gpurun_out/${tag}_hot -f $B --profile-iso 1 > gpurun_out/${tag}_ncu_b.log 2>&1; echo "ncu hot rc=$?"
for f in knn hot; do
  [ -f gpurun_out/${tag}_$f.ncu-rep ] && ncu -i gpurun_out/${tag}_$f.ncu-rep --page raw --csv > gpurun_out/${tag}_${f}_raw.csv 2>/dev/null
done
# gpurun brings back at most 64 MiB: keep the CSV pages, drop the reports
[ -f gpurun_out/${tag}_knn.ncu-rep ] && ncu -i gpurun_out/${tag}_knn.ncu-rep --page source --csv --kernel-name regex:k_map_knn --launch-skip 0 --launch-count 1 > gpurun_out/${tag}_knn_source_iter0.csv 2>/dev/null
rm -f gpurun_out/${tag}_*.ncu-rep
ls -la gpurun_out/ | grep $tag
