"""bench.py contract checks that need no GPU: the reference arm's JSON line, rank gating, and the loud failure of our arm."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ); e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--workload", "cornell", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Msamples/s at 1920x1080" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "1920x1080" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_runs_on_rank_0_only():
    r = _run(["--impl", "reference", "--workload", "cornell", "--steps", "1", "--warmup", "1", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="GPU present")
def test_our_arm_fails_loudly_without_a_gpu():
    r = _run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
