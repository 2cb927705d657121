#!/bin/bash
# round 2, last build, part C: ncu launch lists, ncu --set full of both kernels of the step per env family.  The reports are
# read on the box (raw page, per-line stall / instruction aggregation) and only the summaries travel back (copy-back limit).
# Numbers printed under ncu are never bench values.
T=gpurun_out/r02fin
mkdir -p $T
summ() {  # $1 = report name (without .ncu-rep)
  ncu -i $T/$1.ncu-rep --page raw --csv > $T/ncu_$1_raw.csv 2>/dev/null
  ncu -i $T/$1.ncu-rep --page source --csv --print-source cuda,sass > $T/$1_src.csv 2>/dev/null
  python scripts/ncu_line_stalls.py $T/$1_src.csv 45 > $T/ncu_$1_lines.txt 2>&1
  rm -f $T/$1_src.csv
  [ "$2" = keep ] || rm -f $T/$1.ncu-rep
}
for v in base eco stag; do
  e=4096; [ $v = eco ] && e=16384; [ $v = stag ] && e=8192
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 120 --csv --log-file $T/launches_$v.csv python bench.py --variant $v --envs $e --groups 1 --steps 40 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_l_$v.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_obs -s 310 -c 1 -f -o $T/obs_$v python bench.py --variant $v --envs $e --groups 1 --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_obs_$v.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_$v -s 310 -c 1 -f -o $T/step_$v python bench.py --variant $v --envs $e --groups 1 --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_step_$v.log 2>&1
  k=drop; [ $v = base ] && k=keep
  summ obs_$v $k; summ step_$v $k
  echo "ncu $v done"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_eco -s 310 -c 1 -f -o $T/step_metabolic python bench.py --variant metabolic --envs 16384 --groups 1 --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_step_metabolic.log 2>&1
summ step_metabolic drop
# the observation kernel at the default run's population (roofline.traffic of the bench line)
timeout 400 ncu --set full --clock-control none -k regex:ppg_obs -s 760 -c 1 -f -o $T/obs_base_default python bench.py --steps 460 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_obs_base_default.log 2>&1
ncu -i $T/obs_base_default.ncu-rep --page raw --csv > $T/ncu_obs_base_default_raw.csv 2>/dev/null; rm -f $T/obs_base_default.ncu-rep
du -sh $T; ls $T
