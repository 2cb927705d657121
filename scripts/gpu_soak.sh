#!/bin/bash
# long rollouts: no hang, no error flag, status_envs 0 after tens of thousands of launches
b() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$name.json")); r=d["roofline"]
    print("$name", "steps=%d value=%.4g ms/step=%.4f status=%s launches=%d"%(d["steps"], d["value"], d["ms_per_step"], d["status_envs"], d["gpu_launches"]))
except Exception as e: print("$name ERR", e, open("gpurun_out/$name.err").read()[-800:])
PY
}
b soak_base python bench.py --steps 30000 --warmup 200 --no-cpu --no-e2e
b soak_add python bench.py --reward-mode additive --envs 16384 --steps 6000 --no-cpu --no-e2e
b soak_eco python bench.py --variant eco --envs 16384 --steps 8000 --no-cpu --no-e2e
b soak_stag python bench.py --variant stag --envs 8192 --steps 5000 --warmup 600 --no-cpu --no-e2e
b soak_seasonal python bench.py --steps 5000 --no-cpu --no-e2e --seasonal
