#!/bin/bash
mkdir -p gpurun_out/r02ab
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_stag.py -m gpu -x -q > gpurun_out/r02ab/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02ab/pytest.log
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --steps 300 --warmup 50 --no-cpu --no-e2e --no-configs > gpurun_out/r02ab/$name.json 2> gpurun_out/r02ab/$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02ab/$name.json")); r=d["roofline"]
    print("$name value %.3e ms/step %.4f obs_ms %.4f obs_frac %.3f step_ms %.4f whole_frac %.3f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["step_kernel_ms"], r["whole_step"]["frac"]))
except Exception as e: print("$name failed", e)
PY
}
run base X=1 -- --variant base --envs 4096
run base_dyn PPG_STATIC_FIRST=0 -- --variant base --envs 4096
run add X=1 -- --variant base --reward-mode additive --envs 16384
run add_dyn PPG_STATIC_FIRST=0 -- --variant base --reward-mode additive --envs 16384
run stag X=1 -- --variant stag --envs 8192
run stag_dyn PPG_STATIC_FIRST=0 -- --variant stag --envs 8192
