#!/bin/bash
# parity subset + phase profile + short benches of a step-kernel change
T=${1:-r02h}
mkdir -p gpurun_out/$T
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_eco.py tests/test_gpu_parity_stag.py tests/test_gpu_parity_traits.py -m gpu -x -q > gpurun_out/$T/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/$T/pytest.log
for v in base eco stag; do
  e=4096; [ $v = eco ] && e=16384; [ $v = stag ] && e=8192
  PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_prof.so python scripts/phase_profile.py --variant $v --envs $e > gpurun_out/$T/phase_$v.txt 2>&1; cat gpurun_out/$T/phase_$v.txt
  python bench.py --variant $v --envs $e --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > gpurun_out/$T/bench_$v.json 2> gpurun_out/$T/bench_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/$T/bench_$v.json")); r=d["roofline"]
print("$v value %.3e ms/step %.4f obs_ms %.4f step_ms %.4f whole_frac %.3f obs_frac %.3f live %.1f status %s"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["step_kernel_ms"], r["whole_step"]["frac"], r["frac"], d["mean_live_agents_per_env"], d.get("status_envs")))
PY
done
python bench.py --variant base --reward-mode additive --envs 16384 --steps 100 --warmup 20 --no-cpu --no-e2e --no-configs > gpurun_out/$T/bench_add.json 2> gpurun_out/$T/bench_add.err
python - <<PY
import json
d=json.load(open("gpurun_out/$T/bench_add.json")); r=d["roofline"]
print("add value %.3e ms/step %.4f obs_ms %.4f step_ms %.4f whole_frac %.3f obs_frac %.3f live %.1f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["step_kernel_ms"], r["whole_step"]["frac"], r["frac"], d["mean_live_agents_per_env"]))
PY
