#!/bin/bash
mkdir -p gpurun_out/r02s
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > gpurun_out/r02s/$name.json 2> gpurun_out/r02s/$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02s/$name.json")); r=d["roofline"]
    print("$name value %.3e ms/step %.4f obs_ms %.4f step_ms %.4f whole_frac %.3f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["step_kernel_ms"], r["whole_step"]["frac"]))
except Exception as e: print("$name failed", e)
PY
}
L=$PWD/predpreygrass_b200
for lib in "" _e0 _d0 _e0d0; do
  for ord in 1 0; do
    run base${lib}_o$ord PPG_LIB=$L/libppg_b200$lib.so PPG_ENV_ORDER=$ord -- --variant base --envs 4096
    run stag${lib}_o$ord PPG_LIB=$L/libppg_b200$lib.so PPG_ENV_ORDER=$ord -- --variant stag --envs 8192
    run add${lib}_o$ord PPG_LIB=$L/libppg_b200$lib.so PPG_ENV_ORDER=$ord -- --variant base --reward-mode additive --envs 16384
  done
  run eco${lib} PPG_LIB=$L/libppg_b200$lib.so -- --variant eco --envs 16384
done
