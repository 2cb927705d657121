#!/bin/bash
T=${1:-full}
mkdir -p gpurun_out/$T
python -m pytest tests -m gpu -q > gpurun_out/$T/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/$T/pytest_gpu.log
