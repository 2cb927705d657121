#!/bin/bash
T=gpurun_out/r02conn
mkdir -p $T
python -m pytest tests/test_gpu_connector.py -m gpu -q -x > $T/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $T/pytest_gpu.log
