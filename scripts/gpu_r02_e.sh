#!/bin/bash
mkdir -p gpurun_out/r02e
python -m pytest tests/test_gpu_parity_eco.py tests/test_gpu_dict_adapters.py tests/test_gpu_fullsize.py tests/test_gpu_pow.py tests/test_gpu_parity_stag.py -m gpu -q > gpurun_out/r02e/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02e/pytest_gpu.log
