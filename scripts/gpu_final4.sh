#!/bin/bash
# final pass of the round: full GPU suite, bench lines, ncu captures of the final observation / step kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
b() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$name.json")); r=d.get("roofline") or {}
    print("$name", "value=%.4g ms/step=%.4f obs_ms=%s step_ms=%s frac=%s status=%s cpu=%s e2e=%s"%(d["value"], d["ms_per_step"], r.get("kernel_ms"), r.get("step_kernel_ms"), r.get("frac"), d.get("status_envs"), (d.get("cpu_baseline") or {}).get("value"), (d.get("e2e") or {}).get("value")))
except Exception as e: print("$name ERR", e, open("gpurun_out/$name.err").read()[-500:])
PY
}
b f4_base python bench.py
b f4_ref python bench.py --impl reference --steps 20 --warmup 3
b f4_eco_rich python bench.py --variant eco --eco-rich --envs 16384 --no-cpu --no-e2e
for v in base stag; do
  envs=4096; [ $v == stag ] && envs=8192
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches4_$v.csv python bench.py --variant $v --envs $envs --steps 40 --warmup 20 --no-cpu --no-e2e > gpurun_out/ncu_l4_$v.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_obs -s 250 -c 2 -o gpurun_out/obs4_$v python bench.py --variant $v --envs $envs --steps 40 --warmup 250 --no-cpu --no-e2e > gpurun_out/ncu_obs4_$v.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppg_step_base -s 250 -c 2 -o gpurun_out/step4_base python bench.py --steps 40 --warmup 250 --no-cpu --no-e2e > gpurun_out/ncu_step4_base.log 2>&1
ls gpurun_out | grep 4
