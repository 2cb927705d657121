#!/bin/bash
# exporters on the device + ncu capture of the observation kernel after the sticky-constants change
T=gpurun_out/r02ak
mkdir -p $T
python -m pytest tests/test_event_log.py tests/test_gpu_dict_adapters.py -m gpu -x -q > $T/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $T/pytest.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_obs -s 310 -c 1 -f -o $T/obs_base python bench.py --variant base --envs 4096 --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_obs_base.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_obs -s 310 -c 1 -f -o $T/obs_stag python bench.py --variant stag --envs 8192 --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_obs_stag.log 2>&1
ls -la $T
