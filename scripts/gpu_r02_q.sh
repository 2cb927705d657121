#!/bin/bash
mkdir -p gpurun_out/r02q
python -m pytest tests/test_gpu_parity_eco.py tests/test_gpu_parity_traits.py -m gpu -x -q > gpurun_out/r02q/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02q/pytest.log
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > gpurun_out/r02q/$name.json 2> gpurun_out/r02q/$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02q/$name.json")); r=d["roofline"]
    print("$name value %.3e ms/step %.4f obs_ms %.4f step_ms %.4f whole_frac %.3f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["step_kernel_ms"], r["whole_step"]["frac"]))
except Exception as e: print("$name failed", e)
PY
}
run eco X=1 -- --variant eco --envs 16384
run metabolic X=1 -- --variant metabolic --envs 16384
run base_c24 PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_c24.so -- --variant base --envs 4096
run add_c24 PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_c24.so -- --variant base --reward-mode additive --envs 16384
run base X=1 -- --variant base --envs 4096
