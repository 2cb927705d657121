#!/bin/bash
# the rest of the GPU suite after the fixed counter test (the first 210 tests passed in r02an)
T=gpurun_out/r02ao
mkdir -p $T
python -m pytest tests/test_gpu_parity_traits.py tests/test_gpu_pow.py -m gpu -q > $T/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $T/pytest_gpu.log
