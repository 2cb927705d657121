#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, one rank per GPU) + the reference arm + smoke()
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L

timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 500 --warmup 200 > gpurun_out/multi_${N}.json 2> gpurun_out/multi_${N}.err; echo "rc=$?"; tail -3 gpurun_out/multi_${N}.err; cat gpurun_out/multi_${N}.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 3 > gpurun_out/multi_ref_${N}.json 2> gpurun_out/multi_ref_${N}.err; echo "rc=$?"; cat gpurun_out/multi_ref_${N}.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --variant stag --envs 8192 --steps 300 --warmup 600 --no-e2e > gpurun_out/multi_stag_${N}.json 2> gpurun_out/multi_stag_${N}.err; echo "rc=$?"; cat gpurun_out/multi_stag_${N}.json
