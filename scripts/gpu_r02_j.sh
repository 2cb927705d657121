#!/bin/bash
mkdir -p gpurun_out/r02j
export PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_prof.so
PPG_OBS_OVERLAP=0 python scripts/phase_profile.py --variant base --envs 4096 > gpurun_out/r02j/phase_base_noov.txt 2>&1; cat gpurun_out/r02j/phase_base_noov.txt
unset PPG_LIB
PPG_OBS_OVERLAP=0 python bench.py --variant base --envs 4096 --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('noov base value %.3e ms/step %.4f obs_ms %.4f step_ms %.4f'%(d['value'], d['ms_per_step'], r['kernel_ms'], r['step_kernel_ms']))"
python scripts/env_cycles.py --variant base --envs 4096 --warmup 300 2>&1 | tail -12
