"""bench.py's reference arm (runs on the CPU) prints exactly one JSON line with the keys the driver reads; the GPU arm
refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _run(*args, timeout=300):
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.parametrize("workload", ["config2", "config4"])
def test_reference_arm_json_contract(workload):
    res = _run("--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0", "--gpus", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout                      # ONE line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("waveform samples/sec") and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and workload in d["config"]["workload"]
    cb = d["cpu_baseline"]
    # the reference's own functions when the staged reference (oracle/_ref, build container / GPU box) is present, else the restatement
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "utterances" in cb["sample"]
    assert d["cpu_port"]["kind"] == "port" and d["cpu_port"]["value"] > 0          # second figure: the C port
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    env_rank = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--workload", "config2"], capture_output=True, text=True, timeout=120, env=env_rank)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = _run("--steps", "1", "--warmup", "0", timeout=120)
    assert res.returncode != 0 and "needs a GPU" in (res.stderr + res.stdout)
