#!/bin/bash
# one ncu --set full capture (with source) of the step kernel of a variant at the steady-state population
v=${1:-base}; envs=${2:-4096}; tag=${3:-r02}
mkdir -p gpurun_out/$tag
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppg_step_$([ $v = stag ] && echo stag || ([ $v = base ] && echo base || echo eco)) -s 310 -c 1 -f -o gpurun_out/$tag/step_$v \
  python bench.py --variant $v --envs $envs --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > gpurun_out/$tag/ncu_step_$v.log 2>&1
tail -3 gpurun_out/$tag/ncu_step_$v.log
ls -la gpurun_out/$tag/
