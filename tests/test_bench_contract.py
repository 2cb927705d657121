"""`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) prints one well-formed JSON line; the
GPU arm refuses to run without a GPU instead of falling back (`-m "not gpu"`)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, env=e, timeout=300)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-envs", "32")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent_steps_per_s_incl_obs" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-envs", "8", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "needs a GPU" in r.stderr
