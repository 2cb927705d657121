#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
b() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$name.json")); r=d["roofline"]
    print("$name", "value=%.4g ms/step=%.4f obs_ms=%.4f step_ms=%.4f frac=%.3f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r.get("step_kernel_ms",0), r["frac"]))
except Exception as e: print("$name ERR", e, open("gpurun_out/$name.err").read()[-500:])
PY
}
b ov_base python bench.py --no-cpu --no-e2e
PPG_OBS_OVERLAP=0 b noov_base python bench.py --no-cpu --no-e2e
for sc in 10 12 14; do for oc in 1 2 4; do PPG_STEP_CTAS_PER_SM=$sc PPG_OBS_CTAS_PER_SM=$oc b ov_base_s${sc}_o${oc} python bench.py --no-cpu --no-e2e --steps 300; done; done
PPG_OBS_CTAS_PER_SM=2 b ov_base_o2 python bench.py --no-cpu --no-e2e --steps 300
PPG_OBS_CTAS_PER_SM=4 b ov_base_o4 python bench.py --no-cpu --no-e2e --steps 300
b ov_add16k python bench.py --reward-mode additive --envs 16384 --no-cpu --no-e2e
b ov_eco python bench.py --variant eco --envs 16384 --no-cpu --no-e2e
b ov_stag python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600
PPG_STEP_CTAS_PER_SM=8 PPG_OBS_CTAS_PER_SM=2 b ov_stag_s8_o2 python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600
PPG_STEP_CTAS_PER_SM=12 PPG_OBS_CTAS_PER_SM=2 b ov_eco_s12_o2 python bench.py --variant eco --envs 16384 --no-cpu --no-e2e
