#!/usr/bin/env python
"""bench.py — agent-steps/s (incl. observation generation) of the batched PredPreyGrass env step.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`, for N > 1
launched under torchrun (one rank per GPU).  Prints ONE JSON line on rank 0.

A "step" is one lockstep `step()` of every env instance the rank owns (BASELINE configs[1]:
base_environment, 4096 envs per GPU, uniform random actions from the device-side Philox action
generator, auto-reset).  `value` = agent-steps/s over all ranks with everything resident in HBM;
`e2e` = the same through `ppg_step_host` (actions from pinned host memory, whole row batch copied
back to pinned host memory every step).  `roofline` is the step kernel's algorithmic bytes over its
CUDA-event duration against MEASURED_PEAKS.json.  `cpu_baseline` is the CPU oracle (a C port of the
reference's Python step; the Python reference itself cannot travel to the GPU box) on host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent_steps_per_s_incl_obs"
UNIT = "agent-steps/s"
_REAL_STDOUT = 1
S_AGENT = 42  # bytes of per-agent state + io per agent-step: pos 2 + energy 8 + id 4 (read+write = 28), action 4, reward 4 + flags 2 + id 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="env instances per GPU (BASELINE configs[1])")
    ap.add_argument("--variant", default="base", choices=["base", "eco", "stag"],
                    help="base: BASELINE configs[1]/[2] (base_environment family); eco: configs[3] (eco_evolutionary, speed trait); "
                         "stag: configs[4] (stag_hunt_forward_view_nature_nurture, team capture)")
    ap.add_argument("--eco-rich", action="store_true", help="eco: reproduction-heavy override (thresholds 8/5, grass regrowth 0.3)")
    ap.add_argument("--seasonal", action="store_true", help="base: base_environment_seasonal config (square-wave grass regrowth)")
    ap.add_argument("--reward-mode", default="sparse")
    ap.add_argument("--cap", type=int, nargs=2, default=None)
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--cpu-envs", type=int, default=2048, help="envs of the CPU sample (cpu_baseline and the reference arm)")
    ap.add_argument("--cpu-steps", type=int, default=400, help="steps of the cpu_baseline sample (~10-20 s of host work)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.cap is None:
        # slot capacity per env and species: large enough that no env of the rollout ever fills it (`status_envs` 0 —
        # a full list would suppress births the reference allows); measured maxima: BASE 26 / 83, ECO 13 / 65,
        # ECO reproduction-heavy 222 / 416+, STAG 127 / 416+ (scripts/status_diag.py)
        args.cap = {"base": [32, 128], "eco": [256, 512] if args.eco_rich else [32, 96], "stag": [160, 640]}[args.variant]
    return args


def workload_name(args):
    if args.variant == "stag":
        return (f"stag_hunt_forward_view_nature_nurture default config_env, {args.envs} envs per GPU, uniform random actions "
                "(predators [9, 2], prey 9), auto-reset, Philox facing/trait/capture draws")
    if args.variant == "eco":
        return (f"eco_evolutionary default config_env{' + reproduction-heavy override' if args.eco_rich else ''}, {args.envs} envs per GPU, "
                "uniform random actions (25), auto-reset, Philox trait/mutation draws")
    name = "base_environment_seasonal" if getattr(args, "seasonal", False) else "base_environment"
    return f"{name} default config_env, {args.envs} envs per GPU, uniform random actions, auto-reset, reward={args.reward_mode}"


def build_config(args, **kw):
    from predpreygrass_b200.config import BASE_CONFIG, ECO_CONFIG, STAG_CONFIG, VARIANT_ECO, VARIANT_STAG, make_config

    if args.variant == "stag":
        return make_config(STAG_CONFIG, variant=VARIANT_STAG, cap_live=tuple(args.cap), **kw)
    if args.variant == "eco":
        d = dict(ECO_CONFIG)
        if args.eco_rich:
            d.update(energy_gain_per_step_grass=0.3, prey_creation_energy_threshold=5.0, predator_creation_energy_threshold=8.0,
                     energy_loss_per_step_predator=0.1, n_possible_predators=8000, n_possible_prey=24000)
        return make_config(d, variant=VARIANT_ECO, cap_live=tuple(args.cap), **kw)
    if getattr(args, "seasonal", False):
        from predpreygrass_b200.config import SEASONAL_CONFIG

        return make_config(SEASONAL_CONFIG, reward_mode=args.reward_mode, cap_live=tuple(args.cap), **kw)
    return make_config(BASE_CONFIG, reward_mode=args.reward_mode, cap_live=tuple(args.cap), **kw)


def n_actions(args):
    return 25 if args.variant == "eco" else 9


def action_pools(args, n, rng):
    """host-side random actions (numpy int32) for n rows per species; STAG predators carry the join_hunt bit (include/ppg.h)"""
    import numpy as np

    a0 = rng.integers(0, n_actions(args), size=n, dtype=np.int32)
    a1 = rng.integers(0, n_actions(args), size=n, dtype=np.int32)
    if args.variant == "stag":
        a0 = a0 | (rng.integers(0, 2, size=n, dtype=np.int32) << 8)
    return a0, a1


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_oracle_rate(args, threads, steps, warmup):
    """agent-steps/s of the CPU oracle (C port of the reference step) on `threads` host threads."""
    import numpy as np

    from oracle.oracle import Oracle

    cfg = build_config(args, seed=12345)
    o = Oracle(cfg, args.cpu_envs, threads=threads)
    o.reset()
    rng = np.random.default_rng(0)
    pool0, pool1 = action_pools(args, args.cpu_envs * (args.cap[0] + args.cap[1]) + 16, rng)

    def run(k):
        t = 0.0
        for _ in range(k):
            t0 = time.perf_counter()
            o.step(pool0, pool1)  # any action value is valid for any row; sampling is not timed
            t += time.perf_counter() - t0
        return t

    run(warmup)
    s0 = o.stats()
    dt = run(steps)
    s1 = o.stats()
    o.close()
    agent_steps = int(s1[1] - s0[1])
    env_steps = int(s1[0] - s0[0])
    return agent_steps / dt, env_steps / dt, dt


def main_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rate, env_rate, dt = cpu_oracle_rate(args, threads, args.steps, max(3, args.warmup))
    sample = f"{args.cpu_envs} envs x {args.steps} steps of the same workload, {threads} threads (envs partitioned over threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "reference arm = CPU oracle (C port of the reference's Python step; the Python reference cannot travel to the GPU box)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "env_steps_per_s": env_rate, "gpu_launches": 0,
    }
    emit(line)


def emit(line):
    """the ONE JSON line of the contract, on the process's real stdout"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    args = parse()
    # libraries print to stdout behind Python's back (NCCL: "NCCL version ..."): everything but the JSON line goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return main_reference(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist

    from predpreygrass_b200.batched import BatchedPredPreyGrass
    from predpreygrass_b200.config import STAT_NAMES

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(3, args.warmup), args.steps

    # one Philox key for the whole job; env_index_base makes the trajectories independent of the sharding
    cfg = build_config(args, seed=1000, env_index_base=rank * args.envs)
    env = BatchedPredPreyGrass(cfg, args.envs, device=local_rank)
    env.reset()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rollout(k):
        for _ in range(k):
            a0, a1 = env.random_actions(4242)
            env.step(a0, a1)

    rollout(W)
    stats0 = env.stats_device().clone()
    launches0 = env.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    rollout(K)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = env.launch_count() - launches0
    stats1 = env.stats_device().clone()
    d = (stats1 - stats0).to(torch.float64)  # this rank's env/agent steps, rows, births... in the timed region
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        # the one collective of this path: the small all-reduce of episode/population statistics
        dist.all_reduce(d, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    tot = dict(zip(STAT_NAMES, d.tolist()))
    value = tot["agent_steps"] / (ms_max * 1e-3)
    env_rate = tot["env_steps"] / (ms_max * 1e-3)

    # ---- per-kernel timing: CUDA events recorded by the library on the launching stream around each of the step's kernels
    KR = min(K, 200)
    sk0 = env.stats_device().clone()
    torch.cuda.synchronize()
    env.profile_begin()
    for _ in range(KR):
        a0, a1 = env.random_actions(4242)
        env.step(a0, a1)
    ms_step_k, ms_obs_k, n_prof = env.profile_end()
    torch.cuda.synchronize()
    sk1 = env.stats_device().clone()
    kd = dict(zip(STAT_NAMES, (sk1 - sk0).tolist()))
    split = ms_obs_k > 0.0
    step_ms, obs_ms = ms_step_k / KR, ms_obs_k / KR
    row_bytes = [4 * env.C * cfg.obs_range[s] ** 2 for s in range(2)]
    s_agent = S_AGENT + (20 if args.variant in ("eco", "stag") else 0)  # + trait 8 (read+write 16), age 2+2
    obs_bytes = (kd["rows_pred"] * row_bytes[0] + kd["rows_prey"] * row_bytes[1]) / KR
    state_bytes = (kd["agent_steps"] * s_agent + kd["env_steps"] * (cfg.n_grass * 16 + 64)) / KR
    alg_bytes = obs_bytes + state_bytes  # bytes of one lockstep step of all envs
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    step_kernel = {"eco": "ppg_step_eco_kernel", "stag": "ppg_step_stag_kernel"}.get(args.variant, "ppg_step_base_kernel")
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(f"{args.variant}_{args.envs}", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    if split:
        # dominant kernel = the observation writer (> 90 % of the bytes); its algorithmic bytes are the rows it writes
        k_ms = obs_ms
        achieved = obs_bytes / (obs_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": "ppg_obs_kernel", "kernel_ms": obs_ms, "algorithmic_bytes_per_launch": obs_bytes,
                    "step_kernel": step_kernel, "step_kernel_ms": step_ms,
                    "whole_step": {"ms": step_ms + obs_ms, "algorithmic_bytes": alg_bytes,
                                   "achieved": alg_bytes / ((step_ms + obs_ms) * 1e-3) / 1e9,
                                   "frac": alg_bytes / ((step_ms + obs_ms) * 1e-3) / 1e9 / peak},
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"}
    else:
        k_ms = step_ms
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": step_kernel, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"}

    # ---- e2e through the C-ABI with host buffers (rank-local rate, summed over ranks)
    e2e = None
    if not args.no_e2e:
        host = env.make_host_buffers(pinned=True)
        p0, p1 = action_pools(args, max(env.row_capacity) + 4096, np.random.default_rng(1))
        pool0, pool1 = torch.from_numpy(p0).pin_memory(), torch.from_numpy(p1).pin_memory()
        h2d = d2h = 0
        n0, n1 = env.out.counts()
        e0 = env.stats_device().clone()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            off = (i * 61) % 4096
            host["actions0"][:n0].copy_(pool0[off:off + n0])  # host->pinned staging of this step's inputs
            host["actions1"][:n1].copy_(pool1[off:off + n1])
            h2d += 4 * (n0 + n1)
            n0, n1 = env.step_host(host)
            d2h += n0 * (row_bytes[0] + 13) + n1 * (row_bytes[1] + 13) + 4 * (args.envs + 1) * 4 + args.envs * 14 + 16
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e1 = env.stats_device().clone()
        ed = (e1 - e0).to(torch.float64)
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ed, op=dist.ReduceOp.SUM)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": float(ed[1].item()) / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d // args.e2e_steps,
               "d2h_bytes_per_step": d2h // args.e2e_steps, "steps": args.e2e_steps,
               "api": "ppg_step_host (pinned host actions in, full row batch incl. observations out)"}

    final = env.stats()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        r, er, dt = cpu_oracle_rate(args, threads, args.cpu_steps, 20)
        cpu = {"value": r, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_envs} envs x {args.cpu_steps} steps of the same workload ({dt:.1f} s), oracle/ C port of the reference step, {threads} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "envs_per_gpu": args.envs, "cap_live": args.cap,
                       "l2": "each step writes ~%.0f MB of observation rows (> 126 MB L2) to fresh addresses; no explicit flush" % (alg_bytes / 1e6),
                       "state": "fp64 energies in HBM, fp32 observations/rewards out"},
            "env_steps_per_s": env_rate,
            "mean_live_agents_per_env": tot["agent_steps"] / max(tot["env_steps"], 1.0),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
            "status_envs": final["status_envs"],
        }
        emit(line)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
