#!/bin/bash
mkdir -p gpurun_out/r02w
python -m pytest tests/test_gpu_parity_eco.py tests/test_gpu_parity_traits.py -m gpu -x -q > gpurun_out/r02w/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02w/pytest.log
for v in eco cadence; do
python bench.py --variant $v --envs 16384 --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > gpurun_out/r02w/$v.json 2> gpurun_out/r02w/$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02w/$v.json")); r=d["roofline"]
print("$v value %.3e ms/step %.4f obs_ms %.4f step_ms %.4f whole_frac %.3f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["step_kernel_ms"], r["whole_step"]["frac"]))
PY
done
CS=/usr/local/cuda/bin/compute-sanitizer
for v in eco metabolic cadence; do
  timeout 600 $CS --tool racecheck --print-limit 20 python scripts/sanitize_rollout.py $v 24 128 > gpurun_out/r02w/sanitizer_racecheck_$v.log 2>&1
  echo "racecheck $v rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/r02w/sanitizer_racecheck_$v.log | tail -1)"
done
