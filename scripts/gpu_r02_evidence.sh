#!/bin/bash
# round 2 evidence pass: sanitizers, ncu launch lists, ncu --set full of both kernels of the step per variant, the default
# bench line and the reference arm.  Numbers printed under ncu / the sanitizer are never bench values.
T=gpurun_out/r02ev
mkdir -p $T
CS=/usr/local/cuda/bin/compute-sanitizer
for v in base add eco stag metabolic cooperation cadence; do
  for tool in memcheck racecheck synccheck; do
    timeout 600 $CS --tool $tool --print-limit 20 python scripts/sanitize_rollout.py $v 24 128 > $T/sanitizer_${tool}_$v.log 2>&1
    echo "$tool $v rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $T/sanitizer_${tool}_$v.log | tail -1) | $(grep ' ok ' $T/sanitizer_${tool}_$v.log | tail -1)"
  done
done
for v in base eco stag; do
  e=4096; [ $v = eco ] && e=16384; [ $v = stag ] && e=8192
  k=$v; 
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 120 --csv --log-file $T/launches_$v.csv python bench.py --variant $v --envs $e --steps 40 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_l_$v.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_obs -s 310 -c 1 -f -o $T/obs_$v python bench.py --variant $v --envs $e --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_obs_$v.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_$k -s 310 -c 1 -f -o $T/step_$v python bench.py --variant $v --envs $e --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_step_$v.log 2>&1
  echo "ncu $v done: $(ls $T | grep -c ncu-rep) reports"
done
python bench.py --impl reference --steps 20 --warmup 5 > $T/bench_ref.json 2> $T/bench_ref.err; head -c 600 $T/bench_ref.json; echo
python bench.py > $T/bench_default.json 2> $T/bench_default.err; tail -c 300 $T/bench_default.err; head -c 1500 $T/bench_default.json; echo
