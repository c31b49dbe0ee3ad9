#!/bin/bash
# Round-1 closing validation (under gpurun): GPU parity suite, default bench line, launch list of one steady-state frame
# (128 lanes, one context; the three warm-up frames are skipped with --launch-skip, which costs no replay time).
tag=${1:-s6}
mkdir -p gpurun_out
( time timeout 330 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log
tail -4 gpurun_out/pytest_gpu_$tag.log
( time timeout 240 python bench.py ) > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench_$tag.json
B="python bench.py --lanes 128 --groups 1 --steps 2 --warmup 3 --skip-e2e --no-extras --knn-frames 0 --no-cpu-baseline --no-full-schedule"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-567} -c ${COUNT:-376} --csv --log-file gpurun_out/launches_${tag}_l128.csv $B > gpurun_out/ncu_list_$tag.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches_${tag}_l128.csv
