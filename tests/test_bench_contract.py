"""CPU: the bench.py contract.  The reference arm (the oracle port on the host cores) prints one JSON line with the keys the
driver reads; the GPU arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "audio-seconds/sec" and line["unit"] == "audio-s/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"].startswith("BASELINE configs[0]")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["unit"] == "audio-s/s" and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.parametrize("workload", ["vad", "realtime", "tts"])
def test_reference_arm_other_configs(workload):
    """configs[1] / [2] / [4] have a CPU arm of their own (north_star: every named shape next to the CPU path)."""
    r = _run("--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "audio-s/s" and line["value"] > 0 and line["gpu_launches"] == 0
    assert workload in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "audio-s in" in cb["sample"]


def test_cpu_legs_of_the_other_configs_are_bounded():
    import time

    sys.path.insert(0, ROOT)
    import bench

    for w in bench.OTHER_CPU:
        t0 = time.perf_counter()
        cb = bench.cpu_baseline_other(w, 1)
        assert cb["value"] > 0 and cb["cores"] == 1 and time.perf_counter() - t0 < 60.0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, PYTHONPATH=ROOT, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu")
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)
