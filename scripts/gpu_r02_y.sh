#!/bin/bash
mkdir -p gpurun_out/r02y
python -m pytest tests/test_gpu_parity_stag.py tests/test_gpu_dict_adapters.py tests/test_abi.py -m gpu -q -k "stag or abi" > gpurun_out/r02y/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02y/pytest.log
python bench.py --variant stag --envs 8192 --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > gpurun_out/r02y/stag.json 2> gpurun_out/r02y/stag.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02y/stag.json")); r=d["roofline"]
print("stag value %.3e ms/step %.4f obs_ms %.4f step_ms %.4f whole_frac %.3f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["step_kernel_ms"], r["whole_step"]["frac"]))
PY
