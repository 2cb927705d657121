#!/bin/bash
# final pass: full GPU test suite + the bench lines of every BASELINE config with the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
b() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$name.json")); r=d.get("roofline") or {}
    print("$name", "value=%.4g ms/step=%.4f obs_ms=%s step_ms=%s frac=%s status=%s cpu=%s e2e=%s"%(d["value"], d["ms_per_step"], r.get("kernel_ms"), r.get("step_kernel_ms"), r.get("frac"), d.get("status_envs"), (d.get("cpu_baseline") or {}).get("value"), (d.get("e2e") or {}).get("value")))
except Exception as e: print("$name ERR", e, open("gpurun_out/$name.err").read()[-500:])
PY
}
b f_base python bench.py
b f_ref python bench.py --impl reference --steps 5 --warmup 3
b f_add python bench.py --reward-mode additive --envs 16384 --no-e2e
b f_eco python bench.py --variant eco --envs 16384 --no-e2e
b f_eco_rich python bench.py --variant eco --eco-rich --envs 16384 --no-cpu --no-e2e
b f_stag python bench.py --variant stag --envs 8192 --no-e2e --warmup 600
b f_stag_small_caps python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600 --cap 64 192
