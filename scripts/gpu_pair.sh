#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
b() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$name.json")); r=d["roofline"]
    print("$name", "value=%.4g ms/step=%.4f obs_ms=%.4f step_ms=%.4f frac=%.3f status=%s"%(d["value"], d["ms_per_step"], r["kernel_ms"], r.get("step_kernel_ms",0), r["frac"], d["status_envs"]))
except Exception as e: print("$name ERR", e, open("gpurun_out/$name.err").read()[-800:])
PY
}
b p24_base python bench.py --no-cpu --no-e2e
b p24_add python bench.py --reward-mode additive --envs 16384 --no-cpu --no-e2e
b p24_eco python bench.py --variant eco --envs 16384 --no-cpu --no-e2e
b p24_stag python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600
for t in 28 20; do
PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_t$t.so b p${t}_base python bench.py --no-cpu --no-e2e
PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_t$t.so b p${t}_add python bench.py --reward-mode additive --envs 16384 --no-cpu --no-e2e
done
