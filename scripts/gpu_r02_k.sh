#!/bin/bash
mkdir -p gpurun_out/r02k
export PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_prof.so
python scripts/phase_profile.py --variant base --envs 4096 > gpurun_out/r02k/phase_base.txt 2>&1; cat gpurun_out/r02k/phase_base.txt
