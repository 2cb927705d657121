#!/bin/bash
# last pass of round 2 on one GPU (ABI-10 build with the trait variants' event counters and host-replay exporters):
# reference arm + default bench line, 3 / 4 env groups for the 16384-env configs, then smoke and the whole GPU suite
T=gpurun_out/r02last
mkdir -p $T
python bench.py --impl reference --steps 20 --warmup 5 > $T/bench_ref.json 2> $T/bench_ref.err
python bench.py > $T/bench_default.json 2> $T/bench_default.err; tail -c 300 $T/bench_default.err
run() { name=$1; shift
  python bench.py "$@" --steps 200 --warmup 30 --no-cpu --no-e2e --no-configs > $T/$name.json 2> $T/$name.err
}
for g in 2 3 4; do run eco_g$g --variant eco --envs 16384 --groups $g; done
for g in 2 3 4; do run add_g$g --variant base --reward-mode additive --envs 16384 --groups $g; done
python - <<PY
import json, glob
for f in sorted(glob.glob("$T/*.json")):
    try:
        d=json.load(open(f)); r=d.get("roofline") or {}
        print(f.split("/")[-1], "value %.3e ms/step %.4f"%(d["value"], d["ms_per_step"]), ("obs %.4f frac %.3f step %.4f whole %.3f live %.1f g%s"%(r.get("kernel_ms",0), r.get("frac",0), r.get("step_kernel_ms",0), (r.get("whole_step") or {}).get("frac",0), d.get("mean_live_agents_per_env",0), d.get("groups"))) if r else "", "e2e %.3e"%d["e2e"]["value"] if d.get("e2e") else "")
    except Exception as e: print(f, "failed", e)
PY
python -c "import __graft_entry__ as g; g.smoke()" > $T/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $T/smoke.log
python -m pytest tests -m gpu -q > $T/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $T/pytest_gpu.log
