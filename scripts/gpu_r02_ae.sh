#!/bin/bash
# ECO step kernel: what do the exact pow / the inlined speed-plane divisions cost (instruction-cache footprint experiments)
T=gpurun_out/r02ae
mkdir -p $T
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > $T/$name.json 2> $T/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$T/$name.json")); r=d["roofline"]
    print("$name value %.3e ms/step %.4f obs_ms %.4f obs_frac %.3f step_ms %.4f whole_frac %.3f live %.1f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["step_kernel_ms"], r["whole_step"]["frac"], d["mean_live_agents_per_env"]))
except Exception as e: print("$name failed", e)
PY
}
run eco X=1 -- --variant eco --envs 16384
for v in fastpow planeni both; do run eco_$v PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_$v.so -- --variant eco --envs 16384; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_eco -s 310 -c 1 -f -o $T/step_eco python bench.py --variant eco --envs 16384 --steps 10 --warmup 5 --no-cpu --no-e2e --no-configs > $T/ncu_step_eco.log 2>&1
ls -la $T | tail -3
