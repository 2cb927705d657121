#!/bin/bash
mkdir -p gpurun_out/r02p
run() { # name, env settings..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > gpurun_out/r02p/$name.json 2> gpurun_out/r02p/$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02p/$name.json")); r=d["roofline"]
    print("$name value %.3e ms/step %.4f obs_ms %.4f step_ms %.4f whole_frac %.3f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["step_kernel_ms"], r["whole_step"]["frac"]))
except Exception as e: print("$name failed", e)
PY
}
for c in 12 14 16 18 20; do
  run base_c${c} PPG_STEP_CTAS_PER_SM=$c -- --variant base --envs 4096
  run base_c${c}_noord PPG_STEP_CTAS_PER_SM=$c PPG_ENV_ORDER=0 -- --variant base --envs 4096
done
for c in 14 16 20; do
  run add_c${c} PPG_STEP_CTAS_PER_SM=$c -- --variant base --reward-mode additive --envs 16384
done
for c in 10 12 16; do
  run stag_c${c} PPG_STEP_CTAS_PER_SM=$c -- --variant stag --envs 8192
  run eco_c${c} PPG_STEP_CTAS_PER_SM=$c -- --variant eco --envs 16384
done
