#!/bin/bash
# the measured path against the oracle (launch chain off by default), then the default workload with the chain off / on
T=gpurun_out/r02roll
mkdir -p $T
python -m pytest tests/test_gpu_rollout.py -m gpu -q --tb=line -p no:faulthandler -rx > $T/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $T/pytest_gpu.log
python bench.py --steps 500 --warmup 200 --no-cpu --no-e2e --no-configs > $T/bench_chain_off.json 2> $T/bench_chain_off.err
PPG_PDL_CHAIN=1 python bench.py --steps 500 --warmup 200 --no-cpu --no-e2e --no-configs > $T/bench_chain_on.json 2> $T/bench_chain_on.err
python - <<PY
import json
for n in ("off","on"):
    d=json.load(open("$T/bench_chain_%s.json"%n)); r=d["roofline"]
    print("chain", n, "value %.4e ms/step %.4f obs %.4f frac %.3f step %.4f whole %.3f live %.2f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["step_kernel_ms"], r["whole_step"]["frac"], d["mean_live_agents_per_env"]))
PY
