#!/bin/bash
# round 2: GPU test suite on the current build + compute-sanitizer passes, logs under gpurun_out/r02c
mkdir -p gpurun_out/r02c
python -m pytest tests -m gpu -x -q > gpurun_out/r02c/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02c/pytest_gpu.log
CS=/usr/local/cuda/bin/compute-sanitizer
for v in base add eco stag; do
  for tool in memcheck racecheck synccheck; do
    timeout 600 $CS --tool $tool --print-limit 20 python scripts/sanitize_rollout.py $v 24 128 > gpurun_out/r02c/sanitizer_${tool}_$v.log 2>&1
    echo "$tool $v rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02c/sanitizer_${tool}_$v.log | tail -1) | $(grep ' ok ' gpurun_out/r02c/sanitizer_${tool}_$v.log | tail -1)"
  done
done
