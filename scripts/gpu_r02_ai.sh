#!/bin/bash
# observation kernel with species-sticky gather constants + PDL chain obs -> actions -> step: full GPU suite + A/B bench
T=gpurun_out/r02ai
mkdir -p $T
python -m pytest tests -m gpu -x -q > $T/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $T/pytest.log
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --steps 300 --warmup 30 --no-cpu --no-e2e --no-configs > $T/$name.json 2> $T/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$T/$name.json")); r=d["roofline"]
    print("$name value %.3e ms/step %.4f obs_ms %.4f obs_frac %.3f step_ms %.4f whole_frac %.3f live %.1f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["step_kernel_ms"], r["whole_step"]["frac"], d["mean_live_agents_per_env"]))
except Exception as e: print("$name failed", e)
PY
}
run base_pdl1 PPG_PDL_CHAIN=1 -- --variant base --envs 4096
run base_pdl0 PPG_PDL_CHAIN=0 -- --variant base --envs 4096
run add_pdl1 PPG_PDL_CHAIN=1 -- --variant base --reward-mode additive --envs 16384
run add_pdl0 PPG_PDL_CHAIN=0 -- --variant base --reward-mode additive --envs 16384
run eco_pdl1 PPG_PDL_CHAIN=1 -- --variant eco --envs 16384
run eco_pdl0 PPG_PDL_CHAIN=0 -- --variant eco --envs 16384
run stag_pdl1 PPG_PDL_CHAIN=1 -- --variant stag --envs 8192
run stag_pdl0 PPG_PDL_CHAIN=0 -- --variant stag --envs 8192
run eco_g2 PPG_PDL_CHAIN=1 -- --variant eco --envs 16384 --groups 2
run add_g2 PPG_PDL_CHAIN=1 -- --variant base --reward-mode additive --envs 16384 --groups 2
