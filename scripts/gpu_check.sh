#!/bin/bash
# run on the GPU box: parity tests, then (if green) bench + launch list + one full ncu capture of the step kernel
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/pytest_gpu.log 2>&1
rc=$?
echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit $rc; fi
PPG_OBS_BULK=1 timeout 600 python -m pytest tests -m gpu -x -q --tb=short -k "small_philox or reward_modes or golden" > gpurun_out/pytest_gpu_bulk.log 2>&1
echo "bulk pytest rc=$?" >> gpurun_out/pytest_gpu_bulk.log
tail -3 gpurun_out/pytest_gpu_bulk.log
timeout 300 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err
cat gpurun_out/bench1.json
PPG_OBS_BULK=1 timeout 300 python bench.py --no-cpu --no-e2e > gpurun_out/bench_bulk.json 2> gpurun_out/bench_bulk.err
cat gpurun_out/bench_bulk.json
if [ "$1" == "prof" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 20 --no-cpu --no-e2e > gpurun_out/b_ncu.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_base -s 250 -c 2 -o gpurun_out/prof_step python bench.py --steps 40 --warmup 200 --no-cpu --no-e2e > gpurun_out/b_ncu2.log 2>&1
fi
