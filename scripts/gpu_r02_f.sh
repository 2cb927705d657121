#!/bin/bash
mkdir -p gpurun_out/r02f
python -m pytest tests/test_gpu_parity_traits.py -m gpu -q > gpurun_out/r02f/pytest_traits.log 2>&1; echo "traits rc=$?"; tail -40 gpurun_out/r02f/pytest_traits.log
python -m pytest tests/test_gpu_parity_eco.py tests/test_gpu_dict_adapters.py -m gpu -q > gpurun_out/r02f/pytest_eco.log 2>&1; echo "eco rc=$?"; tail -15 gpurun_out/r02f/pytest_eco.log
