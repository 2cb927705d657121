#!/bin/bash
# run on the GPU box: parity tests, then (if green) bench + launch list + one full ncu capture of the step kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/pytest_gpu.log 2>&1
rc=$?
echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit $rc; fi
timeout 300 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err
cat gpurun_out/bench1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 300 python bench.py --variant eco --envs 16384 --no-cpu > gpurun_out/bench_eco.json 2> gpurun_out/bench_eco.err
cat gpurun_out/bench_eco.json
timeout 300 python bench.py --variant eco --eco-rich --envs 16384 --no-cpu --no-e2e > gpurun_out/bench_eco_rich.json 2> gpurun_out/bench_eco_rich.err
cat gpurun_out/bench_eco_rich.json
timeout 300 python bench.py --reward-mode additive --envs 16384 --no-cpu --no-e2e > gpurun_out/bench_add.json 2> gpurun_out/bench_add.err
cat gpurun_out/bench_add.json
if [ "$1" == "prof" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 20 --no-cpu --no-e2e > gpurun_out/b_ncu.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_base -s 250 -c 2 -o gpurun_out/prof_step python bench.py --steps 40 --warmup 200 --no-cpu --no-e2e > gpurun_out/b_ncu2.log 2>&1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_eco.csv python bench.py --variant eco --envs 16384 --steps 40 --warmup 20 --no-cpu --no-e2e > gpurun_out/b_ncu_eco.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_eco -s 250 -c 2 -o gpurun_out/prof_step_eco python bench.py --variant eco --envs 16384 --steps 40 --warmup 200 --no-cpu --no-e2e > gpurun_out/b_ncu2_eco.log 2>&1
fi
