#!/bin/bash
# round 2, last build, part B: sanitizers, ncu launch lists, ncu --set full of both kernels of the step per env family.
# Numbers printed under ncu / the sanitizer are never bench values.
T=gpurun_out/r02fin
mkdir -p $T
CS=/usr/local/cuda/bin/compute-sanitizer
for v in base eco eco_lean metabolic cadence stag; do
  for tool in memcheck racecheck synccheck; do
    timeout 600 $CS --tool $tool --print-limit 20 python scripts/sanitize_rollout.py $v 24 128 > $T/sanitizer_${tool}_$v.log 2>&1
    echo "$tool $v rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $T/sanitizer_${tool}_$v.log | tail -1) | $(grep ' ok ' $T/sanitizer_${tool}_$v.log | tail -1)"
  done
done
for v in base eco stag; do
  e=4096; [ $v = eco ] && e=16384; [ $v = stag ] && e=8192
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 120 --csv --log-file $T/launches_$v.csv python bench.py --variant $v --envs $e --groups 1 --steps 40 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_l_$v.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_obs -s 310 -c 1 -f -o $T/obs_$v python bench.py --variant $v --envs $e --groups 1 --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_obs_$v.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_$v -s 310 -c 1 -f -o $T/step_$v python bench.py --variant $v --envs $e --groups 1 --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_step_$v.log 2>&1
  echo "ncu $v done: $(ls $T | grep -c ncu-rep) reports"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_eco -s 310 -c 1 -f -o $T/step_metabolic python bench.py --variant metabolic --envs 16384 --groups 1 --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_step_metabolic.log 2>&1
# the observation kernel at the default run's population (roofline.traffic of the bench line)
timeout 400 ncu --set full --clock-control none -k regex:ppg_obs -s 760 -c 1 -f -o $T/obs_base_default python bench.py --steps 460 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_obs_base_default.log 2>&1
ls $T | grep -c ncu-rep
