#!/bin/bash
# compute-sanitizer over the frame path (race / memory checks for kernels that rely on shared and global atomics).
# usage (on the GPU box): bash tools/sanitize.sh <tag>   -> gpurun_out/sanitize_<tag>_{memcheck,racecheck,synccheck,initcheck}.log
TAG=${1:-x}
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/sanitize_${TAG}_${tool}.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok' gpurun_out/sanitize_${TAG}_${tool}.log | tr '\n' ' ')"
done
