#!/bin/bash
mkdir -p gpurun_out/r02o
python scripts/env_cycles.py --variant base --envs 4096 --warmup 300 2>&1 | tail -16 | tee gpurun_out/r02o/env_cycles_base.txt
PPG_ENV_ORDER=0 python scripts/env_cycles.py --variant base --envs 4096 --warmup 300 2>&1 | tail -16 | tee gpurun_out/r02o/env_cycles_base_noorder.txt
