b() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$name.json")); r=d["roofline"]
    print("$name", "value=%.4g ms/step=%.4f obs_ms=%.4f step_ms=%.4f frac=%.3f status=%s live=%.1f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r.get("step_kernel_ms",0), r["frac"], d["status_envs"], d["mean_live_agents_per_env"]))
except Exception as e: print("$name ERR", e, open("gpurun_out/$name.err").read()[-500:])
PY
}
for w in 2 8; do
PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_w$w.so b w${w}_base python bench.py --no-cpu --no-e2e
PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_w$w.so b w${w}_add python bench.py --reward-mode additive --envs 16384 --no-cpu --no-e2e
PPG_LIB=$PWD/predpreygrass_b200/libppg_b200_w$w.so b w${w}_stag python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600
done
b w4_base python bench.py --no-cpu --no-e2e
b w4_stag python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600
