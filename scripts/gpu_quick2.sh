b() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$name.json")); r=d["roofline"]
    print("$name", "value=%.4g ms/step=%.4f obs_ms=%.4f step_ms=%.4f frac=%.3f status=%s"%(d["value"], d["ms_per_step"], r["kernel_ms"], r.get("step_kernel_ms",0), r["frac"], d["status_envs"]))
except Exception as e: print("$name ERR", e, open("gpurun_out/$name.err").read()[-800:])
PY
}
PPG_OBS_WARPS=8 b w8_base python bench.py --no-cpu --no-e2e
PPG_OBS_WARPS=8 b w8_add python bench.py --reward-mode additive --envs 16384 --no-cpu --no-e2e
PPG_OBS_WARPS=8 b w8_eco python bench.py --variant eco --envs 16384 --no-cpu --no-e2e
PPG_OBS_WARPS=4 b w4_stag python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600
PPG_OBS_WARPS=8 b w8_stag_small python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600 --cap 64 192
b w4_stag_small python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600 --cap 64 192
