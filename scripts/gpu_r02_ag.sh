#!/bin/bash
# one step kernel per trait variant (trait_mode a template constant): parity + bench
T=gpurun_out/r02ag
mkdir -p $T
python -m pytest tests/test_gpu_parity_traits.py tests/test_trait_known_answers.py tests/test_gpu_dict_adapters.py -m gpu -x -q > $T/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $T/pytest.log
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
for v in cadence metabolic investment cooperation; do run $v X=1 -- --variant $v --envs 16384; done
for v in metabolic cooperation; do run ${v}_g2 X=1 -- --variant $v --envs 16384 --groups 2; done
run add_g2 X=1 -- --variant base --reward-mode additive --envs 16384 --groups 2
run add_g1 X=1 -- --variant base --reward-mode additive --envs 16384
