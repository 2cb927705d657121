#!/bin/bash
mkdir -p gpurun_out/r02v
python -m pytest tests/test_gpu_parity_eco.py tests/test_gpu_parity.py tests/test_gpu_parity_stag.py tests/test_abi.py tests/test_config_limits.py -m gpu -x -q > gpurun_out/r02v/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02v/pytest.log
run() { name=$1; shift; python bench.py "$@" --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > gpurun_out/r02v/$name.json 2> gpurun_out/r02v/$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02v/$name.json")); r=d["roofline"]
    print("$name value %.3e ms/step %.4f obs_ms %.4f step_ms %.4f whole_frac %.3f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["step_kernel_ms"], r["whole_step"]["frac"]))
except Exception as e: print("$name failed", e)
PY
}
run base --variant base --envs 4096
run add --variant base --reward-mode additive --envs 16384
run eco --variant eco --envs 16384
run stag --variant stag --envs 8192
