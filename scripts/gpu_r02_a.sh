#!/bin/bash
# round 2, first pass: the new bench line (pre-rolled population) and the env-group pipelining experiment
mkdir -p gpurun_out/r02a
for g in 1 2 4; do
  python bench.py --steps 300 --warmup 50 --no-cpu --no-configs --groups $g > gpurun_out/r02a/base_g$g.json 2> gpurun_out/r02a/base_g$g.err
  tail -c 300 gpurun_out/r02a/base_g$g.err
done
for v in eco stag; do
  for g in 1 2; do
    e=16384; [ $v = stag ] && e=8192
    python bench.py --variant $v --envs $e --steps 200 --warmup 20 --no-cpu --no-configs --no-e2e --groups $g > gpurun_out/r02a/${v}_g$g.json 2> gpurun_out/r02a/${v}_g$g.err
    tail -c 300 gpurun_out/r02a/${v}_g$g.err
  done
done
python bench.py --reward-mode additive --envs 16384 --steps 200 --warmup 20 --no-cpu --no-configs --no-e2e --groups 2 > gpurun_out/r02a/add_g2.json 2> gpurun_out/r02a/add_g2.err
python bench.py --reward-mode additive --envs 16384 --steps 200 --warmup 20 --no-cpu --no-configs --no-e2e --groups 1 > gpurun_out/r02a/add_g1.json 2> gpurun_out/r02a/add_g1.err
for f in gpurun_out/r02a/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d["roofline"]
    print({k:d[k] for k in ("value","ms_per_step","mean_live_agents_per_env","groups","gpu_launches")}, "obs_ms",r["kernel_ms"],"step_ms",r["step_kernel_ms"],"frac",r["frac"],"whole",r["whole_step"]["frac"], "e2e", (d.get("e2e") or {}).get("value"), "e2e_dev", (d.get("e2e_device_policy") or {}).get("value"))
except Exception as ex:
    print("ERR", ex)
PY
done
