#!/bin/bash
mkdir -p gpurun_out/r02z
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --steps 300 --warmup 50 --no-cpu --no-e2e --no-configs > gpurun_out/r02z/$name.json 2> gpurun_out/r02z/$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02z/$name.json")); r=d["roofline"]
    print("$name value %.3e ms/step %.4f obs_ms %.4f obs_frac %.3f step_ms %.4f whole_frac %.3f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["step_kernel_ms"], r["whole_step"]["frac"]))
except Exception as e: print("$name failed", e)
PY
}
for c in 3 4 5 6; do run base_obs$c PPG_OBS_CTAS_PER_SM=$c -- --variant base --envs 4096; done
run base_w8 PPG_OBS_WARPS=8 -- --variant base --envs 4096
run base X=1 -- --variant base --envs 4096
