#!/bin/bash
mkdir -p gpurun_out/r02d
python -m pytest tests -m gpu -x -q > gpurun_out/r02d/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02d/pytest_gpu.log
CS=/usr/local/cuda/bin/compute-sanitizer
for v in base eco; do
  timeout 600 $CS --tool racecheck --print-limit 20 python scripts/sanitize_rollout.py $v 24 128 > gpurun_out/r02d/sanitizer_racecheck_$v.log 2>&1
  echo "racecheck $v rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/r02d/sanitizer_racecheck_$v.log | tail -1) | $(grep ' ok ' gpurun_out/r02d/sanitizer_racecheck_$v.log | tail -1)"
done
python bench.py --steps 100 --warmup 10 > gpurun_out/r02d/bench_default.json 2> gpurun_out/r02d/bench_default.err; tail -c 600 gpurun_out/r02d/bench_default.err; head -c 3000 gpurun_out/r02d/bench_default.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02d/bench_ref.json 2> gpurun_out/r02d/bench_ref.err; head -c 1500 gpurun_out/r02d/bench_ref.json
