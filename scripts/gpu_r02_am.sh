#!/bin/bash
# observation kernel: the last rows of a species handed out one by one (PPG_OBS_TAIL rows; default one per warp of the CTA)
T=gpurun_out/r02am
mkdir -p $T
python -m pytest tests/test_gpu_parity_stag.py tests/test_gpu_parity.py tests/test_gpu_parity_eco.py tests/test_gpu_parity_traits.py -m gpu -x -q > $T/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $T/pytest.log
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --steps 300 --warmup 30 --no-cpu --no-e2e --no-configs > $T/$name.json 2> $T/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$T/$name.json")); r=d["roofline"]
    print("$name g%d value %.3e ms/step %.4f obs_ms %.4f obs_frac %.3f step_ms %.4f whole_frac %.3f live %.1f"%(d["groups"], d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["step_kernel_ms"], r["whole_step"]["frac"], d["mean_live_agents_per_env"]))
except Exception as e: print("$name failed", e)
PY
}
for t in 0 2 4 8; do run base_t$t PPG_OBS_TAIL=$t -- --variant base --envs 4096; done
for t in 0 4 8 16; do run stag_t$t PPG_OBS_TAIL=$t -- --variant stag --envs 8192; done
for t in 0 4 8; do run eco_t$t PPG_OBS_TAIL=$t -- --variant eco --envs 16384 --groups 1; done
for t in 0 4; do run add_t$t PPG_OBS_TAIL=$t -- --variant base --reward-mode additive --envs 16384 --groups 1; done
