"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the CPU port of the
path on the host cores and prints ONE JSON line with the keys the driver reads; without a CUDA device our own
arm refuses to run instead of falling back to anything."""
import json
import os
import subprocess
import sys

from util import ROOT


def run_bench(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    out = run_bench("--impl", "reference", "--steps", "2", "--warmup", "3", "--batch", "256", "--hidden", "128", "64", "64", "32")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "transition-updates/sec" and line["unit"] == "transitions/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["steps"] == 2 and line["warmup"] == 3 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "UpdateActorCritic" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    # the config block names the workload only and is what the native arm prints for the same flags
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    ns = argparse.Namespace(state_size=58, batch=256, hidden=[128, 64, 64, 32], replay=1_000_000)
    assert line["config"] == bench.workload_config(ns, 1)
    assert "precision" not in line["config"] and "host cores" in line["implementation"]["precision"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout, out.stderr[-500:])


def test_own_arm_refuses_to_run_without_a_gpu():
    out = run_bench("--steps", "1", "--warmup", "3")
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
