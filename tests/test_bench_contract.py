"""CPU: the driver-facing contract of bench.py that can be checked without a GPU -- the reference arm prints ONE JSON
line with the agreed keys (metric / unit / value / e2e / cpu_baseline / impl), ranks other than 0 print nothing, and the
CUDA arm refuses to run on a box without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_prints_one_json_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--batch-per-gpu", "32", "--cpu-sample", "16",
              "--seconds", "0.25"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sounds/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("4s@44.1kHz sounds/sec") and d["value"] > 0 and d["steps"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "sounds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["config"]["per_gpu_batch"] == 32 and "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    p = _run(["--steps", "1", "--warmup", "0"])
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
