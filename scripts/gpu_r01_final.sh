#!/bin/bash
# round-1 measurement pass on the GPU box: bench lines of every BASELINE config, the reference arm, launch lists and
# full ncu captures of both kernels of the step for the three env variants (numbers under ncu are never bench values)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt
b() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; head -c 2500 gpurun_out/$name.json; echo; }
b bench_base python bench.py
b bench_ref python bench.py --impl reference --steps 5 --warmup 3
b bench_add python bench.py --reward-mode additive --envs 16384 --no-cpu --no-e2e
b bench_eco python bench.py --variant eco --envs 16384 --no-e2e
b bench_eco_rich python bench.py --variant eco --eco-rich --envs 16384 --no-cpu --no-e2e
b bench_stag python bench.py --variant stag --envs 8192 --no-e2e --warmup 600
PPG_OBS_OVERLAP=0 b bench_base_noov python bench.py --no-cpu --no-e2e
for v in base eco stag; do
  envs=16384; [ $v == stag ] && envs=8192; [ $v == base ] && envs=4096
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches_$v.csv python bench.py --variant $v --envs $envs --steps 40 --warmup 20 --no-cpu --no-e2e > gpurun_out/ncu_l_$v.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_obs -s 250 -c 2 -o gpurun_out/obs_$v python bench.py --variant $v --envs $envs --steps 40 --warmup 250 --no-cpu --no-e2e > gpurun_out/ncu_obs_$v.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_$v -s 250 -c 2 -o gpurun_out/step_$v python bench.py --variant $v --envs $envs --steps 40 --warmup 250 --no-cpu --no-e2e > gpurun_out/ncu_step_$v.log 2>&1
done
ls -la gpurun_out
