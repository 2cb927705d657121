#!/bin/bash
mkdir -p gpurun_out/r02x
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > gpurun_out/r02x/$name.json 2> gpurun_out/r02x/$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02x/$name.json")); r=d["roofline"]
    print("$name value %.3e ms/step %.4f obs_ms %.4f step_ms %.4f whole_frac %.3f status %s"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["step_kernel_ms"], r["whole_step"]["frac"], d.get("status_envs")))
except Exception as e: print("$name failed", e)
PY
}
L=$PWD/predpreygrass_b200
for c in 22 23; do
  run base_c$c PPG_LIB=$L/libppg_b200_c$c.so -- --variant base --envs 4096
  run add_c$c PPG_LIB=$L/libppg_b200_c$c.so -- --variant base --reward-mode additive --envs 16384
done
run base X=1 -- --variant base --envs 4096
run add X=1 -- --variant base --reward-mode additive --envs 16384
run stag_cap128_448 X=1 -- --variant stag --envs 8192 --cap 128 448
run stag_cap96_448 X=1 -- --variant stag --envs 8192 --cap 96 448
run stag X=1 -- --variant stag --envs 8192
