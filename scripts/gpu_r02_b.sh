#!/bin/bash
# round 2: env-group pipelining with the launches issued from C (ppg_rollout_random), step-kernel residency sweep
mkdir -p gpurun_out/r02b
run() { # name, env assignments..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --no-cpu --no-configs --no-e2e "$@" > gpurun_out/r02b/$name.json 2> gpurun_out/r02b/$name.err || tail -c 400 gpurun_out/r02b/$name.err
}
for g in 1 2 4; do
  run base_g$g -- --steps 300 --warmup 50 --groups $g
  run base_g${g}_s12 PPG_STEP_CTAS_PER_SM=12 -- --steps 300 --warmup 50 --groups $g
  run base_g${g}_s8 PPG_STEP_CTAS_PER_SM=8 -- --steps 300 --warmup 50 --groups $g
done
run base_g2_o3 PPG_OBS_CTAS_PER_SM=3 -- --steps 300 --warmup 50 --groups 2
run base_g2_s12_o3 PPG_STEP_CTAS_PER_SM=12 PPG_OBS_CTAS_PER_SM=3 -- --steps 300 --warmup 50 --groups 2
for g in 1 2 4; do
  run eco_g$g -- --variant eco --envs 16384 --steps 200 --warmup 20 --groups $g
  run stag_g$g -- --variant stag --envs 8192 --steps 200 --warmup 20 --groups $g
  run add_g$g -- --reward-mode additive --envs 16384 --steps 200 --warmup 20 --groups $g
done
run eco_g2_s10 PPG_STEP_CTAS_PER_SM=10 -- --variant eco --envs 16384 --steps 200 --warmup 20 --groups 2
run stag_g2_s6 PPG_STEP_CTAS_PER_SM=6 -- --variant stag --envs 8192 --steps 200 --warmup 20 --groups 2
for f in gpurun_out/r02b/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d["roofline"]
    print("  val %.3e ms %.4f host_ms %.4f live %.1f | obs_ms %.4f step_ms %.4f frac %.3f whole %.3f" % (d["value"], d["ms_per_step"], d["host_issue_ms_per_step"], d["mean_live_agents_per_env"], r["kernel_ms"], r["step_kernel_ms"], r["frac"], r["whole_step"]["frac"]))
except Exception as ex:
    print("ERR", ex)
PY
done
