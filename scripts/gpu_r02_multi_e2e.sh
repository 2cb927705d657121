#!/bin/bash
# e2e at N GPUs with and without binding every rank to the cores next to its GPU (pinned buffers on the GPU's NUMA node)
N=${1:-8}
T=gpurun_out/r02multi
mkdir -p $T
nvidia-smi topo -m > $T/topo_${N}gpu.txt 2>&1; nproc; lscpu | grep -E "Socket|NUMA node\(s\)|^CPU\(s\)"
for b in 1 0; do
  PPG_BENCH_BIND=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$b bench.py --gpus $N --steps 100 --warmup 10 --no-cpu --no-configs > $T/bench_${N}gpu_bind$b.json 2> $T/bench_${N}gpu_bind$b.err; echo "rc=$?"
  grep "bound to" $T/bench_${N}gpu_bind$b.err | head -3
  python - <<PY
import json
d=json.load(open("$T/bench_${N}gpu_bind$b.json"))
print("bind=$b value %.3e e2e %.3e (%.1f GB/s per rank) device_policy %.3e"%(d["value"], d["e2e"]["value"], d["e2e"]["d2h_gbs_this_rank"], d["e2e_device_policy"]["value"]))
PY
done
