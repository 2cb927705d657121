#!/bin/bash
# run on the GPU box: parity tests, benches of every BASELINE config, launch lists + full ncu captures of the ECO / STAG step kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/pytest_gpu.log 2>&1
rc=$?
echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err; cat gpurun_out/bench_base.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 300 python bench.py --reward-mode additive --envs 16384 --no-cpu --no-e2e > gpurun_out/bench_add.json 2> gpurun_out/bench_add.err; cat gpurun_out/bench_add.json
timeout 300 python bench.py --variant eco --envs 16384 --no-cpu > gpurun_out/bench_eco.json 2> gpurun_out/bench_eco.err; cat gpurun_out/bench_eco.json
timeout 300 python bench.py --variant eco --eco-rich --envs 16384 --no-cpu --no-e2e > gpurun_out/bench_eco_rich.json 2> gpurun_out/bench_eco_rich.err; cat gpurun_out/bench_eco_rich.json
timeout 300 python bench.py --variant stag --envs 8192 --no-cpu --warmup 600 > gpurun_out/bench_stag.json 2> gpurun_out/bench_stag.err; cat gpurun_out/bench_stag.json
if [ "$1" == "prof" ]; then
  for v in eco stag; do
    envs=16384; [ $v == stag ] && envs=8192
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_$v.csv python bench.py --variant $v --envs $envs --steps 40 --warmup 20 --no-cpu --no-e2e > gpurun_out/b_ncu_$v.log 2>&1
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_$v -s 250 -c 2 -o gpurun_out/prof_step_$v python bench.py --variant $v --envs $envs --steps 40 --warmup 250 --no-cpu --no-e2e > gpurun_out/b_ncu2_$v.log 2>&1
  done
fi
