#!/bin/bash
# the host-replay exporters of the trait variants on the device + every dict adapter with the complete metric key sets
T=gpurun_out/r02ap
mkdir -p $T
python -m pytest tests/test_event_log.py tests/test_gpu_dict_adapters.py -m gpu -q > $T/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $T/pytest_gpu.log
