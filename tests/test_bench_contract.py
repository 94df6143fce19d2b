"""bench.py's output contract, on CPU: the reference arm prints one JSON line
with the agreed keys, and the CUDA arm refuses to run without a device."""

import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_prints_one_json_line():
    proc = subprocess.run(
        [sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "1",
         "--ref-digits", "13"], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in rec, key
    assert rec["unit"] == "terms/s" and rec["value"] > 0 and rec["vs_baseline"] is None
    assert rec["cpu_baseline"]["kind"] in ("reference", "port")
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["value"] == rec["value"]
    assert "workload" in rec["config"] and "sample" in rec["config"]


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2"],
                          capture_output=True, text=True, timeout=120, env=env)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_cuda_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    proc = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1", "--n", "12"],
                          capture_output=True, text=True, timeout=300)
    assert proc.returncode != 0
    assert "no CPU fallback" in (proc.stderr + proc.stdout)
