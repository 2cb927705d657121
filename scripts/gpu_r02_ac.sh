#!/bin/bash
mkdir -p gpurun_out/r02ac
PPG_OBS_OVERLAP=0 PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_prof.so python scripts/span_profile.py --variant base --envs 4096 2>&1 | tail -4 | tee gpurun_out/r02ac/span_base.txt
