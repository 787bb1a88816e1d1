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


def test_kernel_table_rooflines_from_a_recorded_pass():
    """kernel_table() on a recorded per-kernel pass (no GPU): traversal kernels are reported against the issue roofline with
    `frac` = the committed capture's issue-slot utilisation, the shading kernel against the HBM peak with algorithmic and
    measured bytes; per-kernel figures come from the serial kernel pass (cnt_k / ms_k), not from the two-lane timed region."""
    sys.path.insert(0, ROOT)
    import bench
    steps, spp, W, H = 6, 16, 1920, 1080
    samples = W * H * spp * steps
    cnt_k = {"samples": samples, "closest_rays": int(2.48 * samples), "shadow_rays": int(1.06 * samples), "closest_hits": int(1.6 * samples), "kernel_launches": 51 * steps}
    tim = {"trace_closest": {"ms": 50.0, "launches": 9 * steps}, "trace_shadow": {"ms": 24.0, "launches": 9 * steps}, "film": {"ms": 1.3, "launches": steps},
           "sort_hits": {"ms": 1.9, "launches": 9 * steps}, "bounce": {"ms": 48.0, "launches": 9 * steps}, "surface": {"ms": 0.0, "launches": 0}}
    m = {"cnt": dict(cnt_k, kernel_launches=2 * 51 * steps), "cnt_k": cnt_k, "ms": 124.0, "ms_k": 127.0, "tim": tim, "clk": {"sm_mhz": 1965.0}}
    kern, roof = bench.kernel_table(m, None, "helmet", "offline", W, H, spp, steps)
    assert set(kern) == {"trace_closest", "trace_shadow", "film", "sort_hits", "shade"}            # fused shading is reported as "shade"
    tc = kern["trace_closest"]
    assert tc["bound"] == "issue" and abs(tc["peak"] - 148 * 4 * 1.965) < 1e-6 and tc["unit"] == "G warp-inst/s"
    prof = bench.measured_profile("helmet", "offline", "k_trace_closest")
    assert prof is not None and abs(tc["frac"] - prof["issue_active_pct"] / 100.0) < 1e-12 and 0.5 < tc["frac"] < 0.9
    assert abs(tc["share_of_step"] - round(50.0 / 127.0, 4)) < 1e-9                                  # shares refer to the kernel pass
    assert abs(tc["grays_per_s"] - cnt_k["closest_rays"] / 50.0e-3 / 1e9) < 1e-9
    assert abs(tc["warp_inst_per_ray_implied_live"] - tc["frac"] * tc["peak"] / tc["grays_per_s"]) < 1e-9
    sh = kern["shade"]
    assert sh["bound"] == "hbm" and sh["bytes_per_unit"] == 424
    assert abs(sh["achieved"] - cnt_k["closest_rays"] * 424 / 48.0e-3 / 1e9) < 1e-6 and abs(sh["frac"] - sh["achieved"] / sh["peak"]) < 1e-12
    assert sh["frac_measured_bytes"] is not None and 0.8 < sh["frac_measured_bytes"] / sh["frac"] < 1.2
    assert roof["kernel"] == "k_trace_closest" and roof["bound"] == "issue" and roof["traffic"] > 0 and "r2am" in roof["traffic_source"]
    rb = bench.rays_block(cnt_k, 127.0)
    assert abs(rb["per_sample"] - 3.54) < 0.01 and abs(rb["gsamples_sbe_per_s"] - (cnt_k["closest_rays"] + cnt_k["shadow_rays"]) / 127.0e-3 / 2e9) < 1e-9
