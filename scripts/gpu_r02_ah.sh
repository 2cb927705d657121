#!/bin/bash
# lean ECO step kernel (KIND 3: no carcass / ghost / juvenile / episode-sum code): parity + bench, on and off
T=gpurun_out/r02ah
mkdir -p $T
python -m pytest tests/test_gpu_parity_eco.py tests/test_eco_known_answers.py tests/test_gpu_fullsize.py tests/test_gpu_dict_adapters.py -m gpu -x -q > $T/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $T/pytest.log
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
run eco_lean X=1 -- --variant eco --envs 16384
run eco_full PPG_ECO_LEAN=0 -- --variant eco --envs 16384
run eco_lean_g2 X=1 -- --variant eco --envs 16384 --groups 2
run eco_full_g2 PPG_ECO_LEAN=0 -- --variant eco --envs 16384 --groups 2
