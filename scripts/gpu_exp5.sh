b() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$name.json")); r=d["roofline"]
    print("$name", "value=%.4g ms/step=%.4f obs_ms=%.4f step_ms=%.4f frac=%.3f status=%s live=%.1f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r.get("step_kernel_ms",0), r["frac"], d["status_envs"], d["mean_live_agents_per_env"]))
except Exception as e: print("$name ERR", e, open("gpurun_out/$name.err").read()[-500:])
PY
}
timeout 120 python scripts/env_cycles.py --warmup 700 2>&1 | tail -42
b t_base python bench.py --no-cpu --no-e2e
b t_stag python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600
PPG_STEP_CTAS_PER_SM=14 b t_base_s14 python bench.py --no-cpu --no-e2e
PPG_STEP_CTAS_PER_SM=12 b t_base_s12 python bench.py --no-cpu --no-e2e
PPG_STEP_CTAS_PER_SM=15 b t_base_s15 python bench.py --no-cpu --no-e2e
