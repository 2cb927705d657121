#!/bin/bash
# N-GPU bench of the last build exactly as the driver launches it (torchrun, one rank per GPU) + the reference arm
N=${1:-2}
T=gpurun_out/r02multi
mkdir -p $T
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $T/bench_${N}gpu.json 2> $T/bench_${N}gpu.err; echo "rc=$?"; tail -3 $T/bench_${N}gpu.err; head -c 2500 $T/bench_${N}gpu.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > $T/bench_ref_${N}gpu.json 2> $T/bench_ref_${N}gpu.err; echo "rc=$?"; head -c 700 $T/bench_ref_${N}gpu.json; echo
if [ "$2" = stag ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --variant stag --envs 8192 --steps 200 --warmup 20 --no-e2e --no-cpu --no-configs > $T/bench_stag_${N}gpu.json 2> $T/bench_stag_${N}gpu.err; echo "rc=$?"; head -c 400 $T/bench_stag_${N}gpu.json; echo
fi
