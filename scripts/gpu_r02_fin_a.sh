#!/bin/bash
# round 2, last build, part A: whole GPU suite, smoke, default bench line + reference arm, one line per config / variant
T=gpurun_out/r02fin
mkdir -p $T
python -m pytest tests -m gpu -q > $T/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $T/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $T/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $T/smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > $T/bench_ref.json 2> $T/bench_ref.err
python bench.py > $T/bench_default.json 2> $T/bench_default.err; tail -c 200 $T/bench_default.err
python bench.py --reward-mode additive --envs 16384 --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > $T/bench_add.json 2> $T/bench_add.err
python bench.py --variant eco --envs 16384 --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > $T/bench_eco.json 2> $T/bench_eco.err
python bench.py --variant eco --envs 16384 --groups 1 --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > $T/bench_eco_g1.json 2> $T/bench_eco_g1.err
python bench.py --variant stag --envs 8192 --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > $T/bench_stag.json 2> $T/bench_stag.err
for v in metabolic investment cooperation cadence; do
  python bench.py --variant $v --envs 16384 --steps 200 --warmup 20 --no-cpu --no-e2e --no-configs > $T/bench_$v.json 2> $T/bench_$v.err
done
python bench.py --variant eco --envs 16384 --eco-rich --cap 256 512 --steps 100 --warmup 20 --no-cpu --no-e2e --no-configs > $T/bench_eco_rich.json 2> $T/bench_eco_rich.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$T/bench_*.json")):
    try:
        d=json.load(open(f)); r=d.get("roofline") or {}
        print(f.split("/")[-1], "g%s value %.3e ms/step %.4f"%(d.get("groups"), d["value"], d["ms_per_step"]), "obs %.4f frac %.3f step %.4f whole %.3f live %.1f status %s"%(r.get("kernel_ms",0), r.get("frac",0), r.get("step_kernel_ms",0), (r.get("whole_step") or {}).get("frac",0), d.get("mean_live_agents_per_env",0), d.get("status_envs")) if r else "")
    except Exception as e: print(f, "failed", e)
PY
