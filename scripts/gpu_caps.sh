b() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$name.json")); r=d["roofline"]
    print("$name", "value=%.4g ms/step=%.4f obs_ms=%.4f step_ms=%.4f frac=%.3f status=%s live=%.1f"%(d["value"], d["ms_per_step"], r["kernel_ms"], r.get("step_kernel_ms",0), r["frac"], d["status_envs"], d["mean_live_agents_per_env"]))
except Exception as e: print("$name ERR", e, open("gpurun_out/$name.err").read()[-500:])
PY
}
b c_stag_128_384 python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600 --cap 128 384
b c_stag_128_512 python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600 --cap 128 512
b c_stag_160_640 python bench.py --variant stag --envs 8192 --no-cpu --no-e2e --warmup 600 --cap 160 640
b c_ecorich_224_416 python bench.py --variant eco --eco-rich --envs 16384 --no-cpu --no-e2e --cap 224 416
b c_ecorich_256_512 python bench.py --variant eco --eco-rich --envs 16384 --no-cpu --no-e2e --cap 256 512
b c_base python bench.py --no-cpu --no-e2e
