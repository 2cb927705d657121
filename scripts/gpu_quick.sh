#!/bin/bash
# quick A/B on the GPU box: a parity subset, then the four bench configs (device-side numbers only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -k "4096 or golden or reproduction_heavy or crowded or fullsize" > gpurun_out/pytest_quick.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_quick.log; tail -4 gpurun_out/pytest_quick.log
b() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$name.json")); r=d["roofline"]
    print("$name", "value=%.4g ms/step=%.4f obs_ms=%.4f step_ms=%.4f frac=%.3f status=%s"%(d["value"], d["ms_per_step"], r["kernel_ms"], r.get("step_kernel_ms",0), r["frac"], d["status_envs"]))
except Exception as e: print("$name ERR", e, open("gpurun_out/$name.err").read()[-800:])
PY
}
b q_base python bench.py --no-cpu --no-e2e
b q_add python bench.py --reward-mode additive --envs 16384 --no-cpu --no-e2e
b q_eco python bench.py --variant eco --envs 16384 --no-cpu --no-e2e
b q_stag python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600
