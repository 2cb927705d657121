#!/bin/bash
mkdir -p gpurun_out/r02aa
PPG_OBS_OVERLAP=0 python scripts/env_cycles.py --variant base --envs 4096 --warmup 300 2>&1 | tail -17 | tee gpurun_out/r02aa/env_cycles_base_noov.txt
