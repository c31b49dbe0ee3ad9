#!/bin/bash
# compute-sanitizer evidence: memcheck / racecheck / initcheck over the GPU parity tests.  usage (under gpurun): profiles/sanitize_r2.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
$CS --version > gpurun_out/sanitizer_${tag}_version.txt 2>&1
echo "== memcheck"; 
timeout 900 $CS --tool memcheck --error-exitcode 99 --print-limit 50 python -m pytest tests -m gpu -x -q -k "extract or ops or odometry" -p no:cacheprovider > gpurun_out/sanitizer_${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -15 gpurun_out/sanitizer_${tag}_memcheck.log
echo "== racecheck";
timeout 700 $CS --tool racecheck --error-exitcode 99 --print-limit 50 python -m pytest tests -m gpu -x -q -k "extract or ops" -p no:cacheprovider > gpurun_out/sanitizer_${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -15 gpurun_out/sanitizer_${tag}_racecheck.log
echo "== initcheck";
timeout 600 $CS --tool initcheck --error-exitcode 99 --print-limit 50 python -m pytest tests -m gpu -x -q -k "extract or ops" -p no:cacheprovider > gpurun_out/sanitizer_${tag}_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -15 gpurun_out/sanitizer_${tag}_initcheck.log
