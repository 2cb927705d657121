#!/bin/bash
# observation kernel: straight-line cut-off rows (STAG), start-up loads overlapped with the first image fetch
T=gpurun_out/r02al
mkdir -p $T
python -m pytest tests/test_gpu_parity_stag.py tests/test_gpu_parity.py tests/test_gpu_parity_eco.py tests/test_gpu_dict_adapters.py tests/test_gpu_parity_traits.py -m gpu -x -q > $T/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $T/pytest.log
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
run base X=1 -- --variant base --envs 4096
run stag X=1 -- --variant stag --envs 8192
run eco_g1 X=1 -- --variant eco --envs 16384 --groups 1
run eco X=1 -- --variant eco --envs 16384
run add X=1 -- --variant base --reward-mode additive --envs 16384
run cadence X=1 -- --variant cadence --envs 16384
run investment X=1 -- --variant investment --envs 16384
run cooperation X=1 -- --variant cooperation --envs 16384
