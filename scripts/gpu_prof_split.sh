#!/bin/bash
# full ncu captures of the step kernels and the observation kernel of the two-kernel step + launch lists
mkdir -p gpurun_out
for v in stag eco base; do
  envs=16384; [ $v == stag ] && envs=8192; [ $v == base ] && envs=4096
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_$v -s 250 -c 2 -o gpurun_out/split_step_$v python bench.py --variant $v --envs $envs --steps 40 --warmup 250 --no-cpu --no-e2e > gpurun_out/ncu_step_$v.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_obs -s 250 -c 2 -o gpurun_out/split_obs_$v python bench.py --variant $v --envs $envs --steps 40 --warmup 250 --no-cpu --no-e2e > gpurun_out/ncu_obs_$v.log 2>&1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches_split_$v.csv python bench.py --variant $v --envs $envs --steps 40 --warmup 20 --no-cpu --no-e2e > gpurun_out/ncu_l_$v.log 2>&1
done
ls -la gpurun_out
