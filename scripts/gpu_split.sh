#!/bin/bash
# two-kernel step: parity first, then A/B benches against the one-kernel step (PPG_OBS_SPLIT=0)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "small or philox or lockstep" > gpurun_out/pytest_quick.log 2>&1
echo "quick rc=$?" >> gpurun_out/pytest_quick.log; tail -5 gpurun_out/pytest_quick.log
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/pytest_gpu.log 2>&1
rc=$?
echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
b() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; tail -c 1500 gpurun_out/$name.json; tail -3 gpurun_out/$name.err; }
b split_base python bench.py --no-cpu --no-e2e
PPG_OBS_SPLIT=0 b fused_base python bench.py --no-cpu --no-e2e
b split_add16k python bench.py --reward-mode additive --envs 16384 --no-cpu --no-e2e
b split_eco python bench.py --variant eco --envs 16384 --no-cpu --no-e2e
b split_eco_rich python bench.py --variant eco --eco-rich --envs 16384 --no-cpu --no-e2e
b split_stag python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600
