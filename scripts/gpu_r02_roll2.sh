#!/bin/bash
# after the fix of the hoisted n_rows load in the action kernels: the measured path with the launch chain on and off
T=gpurun_out/r02roll2
mkdir -p $T
python -m pytest tests/test_gpu_rollout.py -m gpu -q --tb=line -p no:faulthandler -rxX > $T/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -9 $T/pytest_gpu.log
