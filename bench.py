#!/usr/bin/env python
"""bench.py — Msamples/s of the path-tracing hot path at 1920x1080 (BASELINE.json metric).

One "step" = one ohb_render() of `--spp-step` samples per pixel over the whole 1080p frame of the
workload (default: configs[1], the Helmet-class ~50K-triangle scene under an HDRI with env importance
sampling + MIS, offline integrator; 16 steps x 16 spp = its 256 spp).  A *sample* is one full OHAO path
tree.  `value` is device-timed with the scene resident in HBM; `e2e` goes through the
C ABI with host buffers in and the RGBA8 image out every step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload helmet|cornell|synthetic2m]
  python bench.py --impl reference      # the CPU restatement of the reference shaders on host cores

Multi-GPU (torchrun, one rank per GPU): the scene is replicated, rank g renders its own contiguous block
of sample indices in sum mode (weak scaling: per-GPU work fixed), and ONE NCCL reduce of the RGBA32F
accumulation image to rank 0 + resolve closes the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

W, H = 1920, 1080
METRIC = "Msamples/s at 1920x1080"


def make_workload(name: str):
    from ohao_engine_b200 import scenes
    if name == "helmet":
        return scenes.helmet_class(), scenes.helmet_camera(), "helmet-class 50K tris + 5x2048^2 textures + 1024x512 HDRI, env IS + MIS, offline, 4 bounces"
    if name == "cornell":
        return scenes.cornell_box(), scenes.cornell_camera(), "cornell_box 2058 tris, 12 sphere lights, offline, 4 bounces"
    if name == "synthetic2m":
        return scenes.synthetic_2m(), scenes.synthetic_camera(), "synthetic 2M tris, 8 lights + HDRI, offline, 4 bounces"
    raise SystemExit(f"unknown workload {name}")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try: self.p.wait(timeout=5)
        except Exception: self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8: continue
            try: sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError: continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"): reasons.add(name)
        try: os.unlink(self.f.name)
        except OSError: pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_rate(ps, cam, seconds: float, nthreads: int, W: int = W, H: int = H):
    """Msamples/s of the CPU restatement (oracle) on full frames of 1 spp, for about `seconds`."""
    from oracle import oracle_py as O
    osc = O.OracleScene(ps)
    v, p = cam.view(), cam.proj(W, H)
    acc = np.zeros((H, W, 4), np.float32)
    t0 = time.perf_counter(); n = 0
    while True:
        osc.render_offline(v, p, W, H, 1, first_sample=n, history=n, accum=acc, nthreads=nthreads, want_ldr=False)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or n >= 64: break
    return W * H * n / dt / 1e6, n, dt


def run_reference(args, rank: int):
    """The reference arm: the reference's own implementation cannot be built here or on the GPU box (Vulkan RT
    pipeline + GLSL, no Vulkan SDK/ICD/glslc — DESIGN.md), so this times the CPU restatement of its shaders
    (oracle/, kind "port") with every host thread on the same workload; each step = 1 spp over the 1080p frame."""
    if rank != 0:
        return
    from oracle import oracle_py as O
    ps, cam, desc = make_workload(args.workload)
    osc = O.OracleScene(ps)
    cores = os.cpu_count() or 1
    v, p = cam.view(), cam.proj(W, H)
    acc = np.zeros((H, W, 4), np.float32)
    idx = 0
    for _ in range(args.warmup):
        osc.render_offline(v, p, W, H, 1, first_sample=idx, history=idx, accum=acc, nthreads=cores, want_ldr=False); idx += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        osc.render_offline(v, p, W, H, 1, first_sample=idx, history=idx, accum=acc, nthreads=cores, want_ldr=True); idx += 1
    dt = time.perf_counter() - t0
    val = W * H * args.steps / dt / 1e6
    sample = f"{args.steps} steps x 1 spp x 1920x1080 of the same scene (our arm: {args.spp_step} spp per step)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "resolution": [W, H], "spp_per_step": 1, "integrator": "offline"},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


SM_ISSUE_SLOTS = 4          # warp schedulers per SM, one warp instruction per clock each


def measured_profile(workload: str, integrator: str, kernel: str):
    """The committed `ncu --set full` capture of this workload (profiles/traffic.json, tools/ncu_traffic.py): per launch of
    `kernel`, DRAM bytes, executed warp instructions, threads per instruction — and the units (rays / paths) that launch
    processed, so the per-unit figures carry over to the launch sizes of THIS run.  None when no capture exists."""
    try:
        e = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[f"{workload}/{integrator}"][kernel]
    except Exception:
        return None
    u = e.get("units_per_launch")
    if not u:
        return None
    return {"dram_bytes_per_unit": e["bytes_per_launch"] / u,
            "warp_inst_per_unit": (e["inst_per_launch"] / u) if e.get("inst_per_launch") else None,
            "thr_per_inst": e.get("thr_per_inst"), "l2_hit_pct": e.get("l2_hit_pct"), "issue_active_pct": e.get("issue_active_pct"), "source": e.get("source")}


def measure(r, ps, cam, workload, integrator, denoise, W, H, spp, steps, warmup, world=1, rank=0, local=0, dist=None, torch=None, sample_clocks=True, strong=False):
    """Warm-up, the device-timed region, the e2e region; returns the pieces of the JSON line (rank 0 assembles them)."""
    from ohao_engine_b200 import binding as B
    import copy, math as _m
    realtime = integrator == "realtime"
    if realtime:
        stg = r.get_settings(); stg.samples_per_frame = spp
        if denoise == "svgf": stg.denoise_mode = B.DENOISE_ATROUS
        r.set_rt_render_settings(stg)
        base_cam = copy.deepcopy(cam); frame_no = [0]
        radius = _m.hypot(base_cam.position[0], base_cam.position[2]); ang0 = _m.atan2(base_cam.position[2], base_cam.position[0])
    v, p = cam.view(), cam.proj(W, H)
    block = (warmup + steps) * spp                      # sample indices per rank: contiguous block, rank-major
    if world > 1 or strong:
        r.set_accum_mode(True)
    r.set_render_seed(rank * block)
    ldr_img = np.empty((H, W, 4), np.uint8)
    image_no = [0]
    accum_t = None
    if world > 1:
        ptr, nbytes = r.accum_dev_ptr()

        class _Alias:      # zero-copy view of the library's accumulation image for the NCCL reduce
            __cuda_array_interface__ = {"shape": (H * W * 4,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
        accum_t = torch.as_tensor(_Alias(), device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        r.synchronize()
        if torch is not None: torch.cuda.synchronize()

    def step():
        """One pass of the hot path: offline = spp samples per pixel; realtime = one frame of the orbit.
        Strong scaling: one step = one COMPLETE image of spp x world samples per pixel — this rank renders its block of
        sample indices into a cleared sum image, ONE NCCL reduce, resolve and the RGBA8 readback on rank 0, all inside."""
        if strong:
            r.clear_accum(); r.set_render_seed(image_no[0] * spp * world + rank * spp); image_no[0] += 1
            r.render(v, p, spp)
            if world > 1:
                r.synchronize(); dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM); torch.cuda.synchronize()
            if rank == 0:
                r.resolve(); r.get_pixels(ldr_img)
            return
        if not realtime:
            r.render(v, p, spp); return
        a = ang0 + _m.radians(0.5 * frame_no[0]); frame_no[0] += 1
        c = copy.deepcopy(base_cam); c.position = (radius * _m.cos(a), base_cam.position[1], radius * _m.sin(a)); c.yaw = _m.degrees(a) + 180.0
        r.notify_camera_changed()
        r.render_realtime(c.view(), p)

    for _ in range(warmup):
        step()
    barrier()
    # ---- timed region: device-resident ---------------------------------------------------------------
    # Offline: the timed region runs the product default (two lanes on two streams, ohb_api.cu ohb_render), where per-kernel event
    # spans would overlap; the per-kernel table comes from a second pass of the same steps with timing on (serial launches, one lane).
    two_pass = not realtime
    r.reset_counters(); r.enable_timing(not two_pass)
    clocks = ClockSampler(local) if (rank == 0 and sample_clocks) else None
    barrier()
    r.timer_start()
    for _ in range(steps):
        step()
    if world > 1 and not strong:
        r.synchronize()
        dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)       # NCCL over NVLink: one reduce per image
        torch.cuda.synchronize()
        if rank == 0: r.resolve()
    ms = r.timer_stop()
    barrier()
    clk = clocks.stop() if clocks else None
    cnt = r.counters()
    if two_pass:
        r.reset_counters(); r.enable_timing(True)
        barrier(); r.timer_start()
        for _ in range(steps):
            step()
        ms_k = r.timer_stop(); barrier()
        cnt_k = r.counters()
    else:
        ms_k, cnt_k = ms, cnt
    tim = r.timing()
    r.enable_timing(False)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        agg = torch.tensor([cnt["samples"], cnt["closest_rays"], cnt["shadow_rays"], cnt["closest_hits"], cnt["kernel_launches"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        tot = dict(zip(("samples", "closest_rays", "shadow_rays", "closest_hits", "kernel_launches"), [int(x) for x in agg.tolist()]))
    else:
        tot = cnt
    samples_total = W * H * spp * steps * world
    # ---- e2e: through the C ABI with host buffers, H2D of the step's inputs + D2H of its image every step ----
    mat = np.ascontiguousarray(ps.mat_colors, np.float32); lights = np.ascontiguousarray(ps.light_ssbo, np.uint8)
    ldr = np.empty((H, W, 4), np.uint8)
    h2d = mat.nbytes + lights.nbytes + 128; d2h = ldr.nbytes
    if not realtime: r.reset_accumulation()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        r.update_rt_material_params(mat); r.update_rt_light_params(lights)      # updateRTMaterialParams / updateRTLightParams
        step()
        if strong: continue                                                        # reduce + resolve + readback are part of a strong-scaling step
        if world > 1: r.resolve()
        r.get_pixels(ldr)                                                          # getPixelSpan(): blocking readback
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    return dict(ms=ms, cnt=cnt, cnt_k=cnt_k, ms_k=ms_k, tot=tot, tim=tim, clk=clk, samples_total=samples_total, value=samples_total / (ms * 1e-3) / 1e6,
                e2e={"value": samples_total / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h})


def kernel_table(m, st, workload, integrator, W, H, spp, steps, num_sms=148):
    """Per-kernel roofline entries.  Traversal kernels are bound by instruction issue, not by bytes: they
    are reported against the issue roofline — warp instructions per ray from the committed ncu capture of this workload x
    the rays/s measured live, over SMs x 4 schedulers x the SM clock sampled during the timed region — beside their measured
    DRAM bytes per ray.  The shading / film / per-pixel kernels are reported against the measured HBM copy peak with their
    algorithmic bytes per unit and, where a capture exists, their measured DRAM bytes."""
    peak, peak_src = peaks()
    cnt, tim, ms = m["cnt_k"], dict(m["tim"]), m["ms_k"]      # the per-kernel pass (== the timed region for the realtime profile)
    realtime = integrator == "realtime"
    clk_mhz = (m["clk"] or {}).get("sm_mhz") or 1965.0
    issue_peak = num_sms * SM_ISSUE_SLOTS * clk_mhz * 1e6 / 1e9           # G warp-inst/s
    units = {"trace_closest": (cnt["closest_rays"], None), "trace_shadow": (cnt["shadow_rays"], None),
             "surface": (cnt["closest_rays"], 248), "bounce": (cnt["closest_rays"], 304), "film": (cnt["samples"], 36),
             "sort_hits": (cnt["closest_rays"], 9),
             "rt_pixel": (W * H * steps, 352 * spp + 416),     # per pixel: N finished path records + 10 history planes in / 6 out
             "svgf": (W * H * steps, 278)}                      # per pixel: temporal pass 78 B + 5 a-trous iterations x 40 B
    fused = tim.get("surface", {}).get("launches", 1) == 0 and not realtime
    if fused:
        # k_shade (surface + bounce fused, payload in registers): 248 + 304 - 2 x 64 B of payload traffic; reported as "shade"
        units["shade"] = (cnt["closest_rays"], 424); tim["shade"] = tim.pop("bounce")
    kern = {}
    for k, x in tim.items():
        if x["launches"] == 0: continue
        nunits, bpu = units[k]
        avg = x["ms"] / x["launches"]; upl = nunits / x["launches"]
        prof = measured_profile(workload, integrator, {"bounce": "k_bounce_rt", "film": "k_rt_denoise" if realtime else "k_film"}.get(k, "k_" + k))
        e = {"ms": round(x["ms"], 3), "launches": x["launches"], "share_of_step": round(x["ms"] / ms, 4), "units_per_launch": upl, "avg_launch_ms": avg,
             "measured_dram_bytes_per_unit": prof["dram_bytes_per_unit"] if prof else None, "profile_source": prof["source"] if prof else None}
        if k.startswith("trace_"):
            # Issue roofline.  frac = the fraction of issue slots the schedulers used (smsp__issue_active, hardware counter, weighted by
            # instructions over the launches of the committed capture); achieved = frac x peak.  Instruction counts cannot be measured
            # outside a profiler, and the capture runs 4 spp per step (its short launches carry relatively more tail), so its
            # warp-instructions per ray are reported as profiled and the live figure is the one implied by the live ray rate.
            wi = prof["warp_inst_per_unit"] if prof else None
            ia = (prof["issue_active_pct"] / 100.0) if (prof and prof.get("issue_active_pct")) else None
            rate = upl / (avg * 1e-3) / 1e9 if avg > 0 else 0.0                     # Grays/s, measured live (CUDA events)
            e.update({"bound": "issue", "frac": ia, "achieved": (ia * issue_peak) if ia else None, "peak": issue_peak, "unit": "G warp-inst/s",
                      "frac_source": "smsp__issue_active.avg.pct_of_peak_sustained_active of the committed ncu capture",
                      "grays_per_s": rate, "warp_inst_per_ray_profiled": wi, "warp_inst_per_ray_implied_live": (ia * issue_peak / rate) if (ia and rate > 0) else None,
                      "thr_per_inst": prof["thr_per_inst"] if prof else None, "l2_hit_pct": prof["l2_hit_pct"] if prof else None,
                      "dram_gbs": (prof["dram_bytes_per_unit"] * rate) if prof else None})
        else:
            ach = upl * bpu / (avg * 1e-3) / 1e9 if avg > 0 else 0.0
            e.update({"bound": "hbm", "bytes_per_unit": bpu, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                      "frac_measured_bytes": (prof["dram_bytes_per_unit"] * upl / (avg * 1e-3) / 1e9 / peak) if (prof and avg > 0) else None})
        kern[k] = e
    dom = max(kern, key=lambda k: kern[k]["ms"]); dk = kern[dom]
    roof = {"bound": dk["bound"], "kernel": "k_" + dom, "achieved": dk["achieved"], "peak": dk["peak"], "unit": dk["unit"], "frac": dk["frac"],
            "traffic": (dk["measured_dram_bytes_per_unit"] * dk["units_per_launch"]) if dk["measured_dram_bytes_per_unit"] else None,
            "traffic_source": dk["profile_source"], "units_per_launch": dk["units_per_launch"], "avg_launch_ms": dk["avg_launch_ms"], "share_of_step": dk["share_of_step"]}
    if dk["bound"] == "hbm":
        roof.update({"peak_source": peak_src, "bytes_per_unit": dk["bytes_per_unit"], "frac_measured_bytes": dk["frac_measured_bytes"]})
    else:
        roof.update({"peak_source": f"{num_sms} SMs x {SM_ISSUE_SLOTS} schedulers x {clk_mhz:.0f} MHz (SM clock sampled during the timed region)",
                     "frac_source": dk["frac_source"], "grays_per_s": dk["grays_per_s"], "warp_inst_per_ray_profiled": dk["warp_inst_per_ray_profiled"],
                     "warp_inst_per_ray_implied_live": dk["warp_inst_per_ray_implied_live"], "thr_per_inst": dk["thr_per_inst"], "l2_hit_pct": dk["l2_hit_pct"], "dram_gbs": dk["dram_gbs"],
                     "hbm_peak": peak})
    return kern, roof


def rays_block(tot, ms):
    nr = tot["closest_rays"] + tot["shadow_rays"]
    return {"per_sample": nr / max(tot["samples"], 1), "closest_per_sample": tot["closest_rays"] / max(tot["samples"], 1),
            "shadow_per_sample": tot["shadow_rays"] / max(tot["samples"], 1), "mrays_per_s": nr / (ms * 1e-3) / 1e6, "gsamples_sbe_per_s": nr / (ms * 1e-3) / 2e9}


# the other BASELINE.json configurations that fit one GPU, measured after the headline in the same process (bounded:
# a few seconds each): (workload, integrator, W, H, spp per step, steps, warm-up)
EXTRA_WORKLOADS = [("synthetic2m", "offline", 1920, 1080, 16, 4, 3),        # north_star's target scene: >= 3 Gsamples/s SBE
                   ("synthetic2m", "realtime", 1920, 1080, 1, 60, 8),       # configs[2]: 1 spp/frame + ReSTIR GI, frames/s
                   ("cornell", "offline", 512, 512, 64, 8, 3)]              # configs[0] at its stated size: 512x512 x 64 spp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="helmet", choices=["helmet", "cornell", "synthetic2m"])
    ap.add_argument("--spp-step", type=int, default=None, help="samples per pixel per step (default 16 offline, 1 realtime)")
    ap.add_argument("--integrator", default="offline", choices=["offline", "realtime"],
                    help="realtime = config 3: one frame per step (ReSTIR GI temporal + spatial, EMA, a-trous), 0.5 deg/frame orbit, N=1 only")
    ap.add_argument("--denoise", default="none", choices=["none", "svgf"],
                    help="realtime only: svgf = DenoiseMode::Atrous (fresh-sample frames + the SVGF denoiser, SURVEY 8f row 2)")
    ap.add_argument("--width", type=int, default=W); ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bounded CPU-baseline sample (rank 0, N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the `workloads` array (the other single-GPU BASELINE configs)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: every step is one fixed image of --image-spp samples per pixel split over the N GPUs (reduce + resolve + readback inside the step)")
    ap.add_argument("--image-spp", type=int, default=256, help="strong scaling: samples per pixel of the fixed image (configs[1]: 256)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.spp_step is None: args.spp_step = 1 if args.integrator == "realtime" else 16

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to fd 1 while the communicator is created; keep stdout to the one JSON line
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
        finally:
            os.dup2(saved, 1); os.close(saved)

    from ohao_engine_b200 import binding as B
    Wd, Hd = args.width, args.height
    ps, cam, desc = make_workload(args.workload)
    realtime = args.integrator == "realtime"
    if realtime and world > 1:
        raise SystemExit("bench.py: the realtime profile shards by tile + halo only (DESIGN.md §5); its headline is N=1")

    def describe(desc, integrator, denoise):
        if integrator != "realtime": return desc
        d = desc.replace("offline, 4 bounces", "realtime, 2 bounces + ReSTIR GI temporal/spatial + EMA + a-trous, 0.5 deg/frame orbit")
        return d.replace("EMA + a-trous", "a-trous, fresh-sample frames + SVGF denoiser (DenoiseMode::Atrous)") if denoise == "svgf" else d

    r = B.Renderer(Wd, Hd, profile=B.PROFILE_REALTIME if realtime else B.PROFILE_OFFLINE, device=local)
    r.set_scene(ps)
    r.build_accel()                                   # second build: the first carries module load + allocation (bvh_build_ms)
    st = r.accel_stats()
    spp = args.spp_step
    strong = args.scaling == "strong"
    if strong:
        if realtime or args.image_spp % world: raise SystemExit("bench.py: --scaling strong needs the offline integrator and --image-spp divisible by the GPU count")
        spp = args.image_spp // world                       # this rank's share of every image
    m = measure(r, ps, cam, args.workload, args.integrator, args.denoise, Wd, Hd, spp, args.steps, args.warmup, world, rank, local, dist, torch, strong=strong)
    ms, tot = m["ms"], m["tot"]

    if rank == 0:
        kern, roof = kernel_table(m, st, args.workload, args.integrator, Wd, Hd, spp, args.steps, torch.cuda.get_device_properties(local).multi_processor_count)
        out = {
            "metric": METRIC, "value": m["value"], "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": describe(desc, args.integrator, args.denoise), "resolution": [Wd, Hd], "spp_per_step": spp * (world if strong else 1), "spp_per_gpu_per_step": spp, "integrator": args.integrator, "tris": int(st.num_tris),
                       "frames_per_s": (args.steps / (ms * 1e-3)) if realtime else None, "treelet_passes": int(st.treelet_passes),
                       "bvh_nodes": int(st.num_nodes), "bvh_sah": round(float(st.sah_cost), 2), "bvh_build_ms": round(float(st.build_ms), 3),
                       "lanes": None if realtime else os.environ.get("OHB_LANES", "auto (2 while nodes + triangles <= 32 MB, else 1)"),
                       "parallelism": (f"one fixed {args.image_spp}-spp image per step, sample-index blocks over {world} GPU(s), NCCL reduce + resolve + RGBA8 readback inside the step" if strong
                                       else (f"spp-shard x{world}" if world > 1 else "single GPU")),
                       "l2": "path-state working set %.1f GB per step >> 126 MB L2 (no explicit flush)" % (Wd * Hd * min(spp, 16) * (337 if realtime else 273) / 1e9)},
            "rays": rays_block(tot, ms),
            "kernels": kern,
            "kernel_pass": None if realtime else {"ms_per_step": m["ms_k"] / args.steps, "lanes": 1,
                            "note": "per-kernel CUDA-event spans need serial launches: `kernels` / `roofline` come from a second pass of the same steps with "
                                    "ohb_enable_timing on (one lane); `value` is the timed region above with the product default (OHB_LANES=%s: two lanes while nodes + triangles <= 32 MB)" % os.environ.get("OHB_LANES", "auto")},
            "roofline": roof,
            "e2e": m["e2e"],
            "gpu_launches": int(m["cnt"]["kernel_launches"]),
            "clocks": m["clk"],
        }
        if world == 1 and not args.no_workloads and not strong:
            # the other single-GPU configurations of BASELINE.json, same process, same box (each a few seconds)
            del r
            out["workloads"] = []
            for wl, integ, w2, h2, spp2, steps2, warm2 in EXTRA_WORKLOADS:
                if (wl, integ, w2, h2) == (args.workload, args.integrator, Wd, Hd): continue
                ps2, cam2, desc2 = make_workload(wl)
                r2 = B.Renderer(w2, h2, profile=B.PROFILE_REALTIME if integ == "realtime" else B.PROFILE_OFFLINE, device=local)
                r2.set_scene(ps2); r2.build_accel(); st2 = r2.accel_stats()
                m2 = measure(r2, ps2, cam2, wl, integ, "none", w2, h2, spp2, steps2, warm2, torch=torch, sample_clocks=True)
                k2, roof2 = kernel_table(m2, st2, wl, integ, w2, h2, spp2, steps2, torch.cuda.get_device_properties(local).multi_processor_count)
                out["workloads"].append({"workload": describe(desc2, integ, "none"), "name": wl, "integrator": integ, "resolution": [w2, h2], "spp_per_step": spp2, "steps": steps2, "warmup": warm2,
                                         "value": m2["value"], "unit": "Msamples/s", "ms_per_step": m2["ms"] / steps2, "frames_per_s": (steps2 / (m2["ms"] * 1e-3)) if integ == "realtime" else None,
                                         "tris": int(st2.num_tris), "bvh_build_ms": round(float(st2.build_ms), 3), "rays": rays_block(m2["tot"], m2["ms"]), "roofline": roof2,
                                         "kernels": {k: {kk: v[kk] for kk in ("ms", "launches", "share_of_step", "bound", "frac") if kk in v} for k, v in k2.items()},
                                         "kernel_pass_ms_per_step": None if integ == "realtime" else m2["ms_k"] / steps2,
                                         "e2e": m2["e2e"], "gpu_launches": int(m2["cnt"]["kernel_launches"]), "clocks": m2["clk"]})
                del r2
                if integ == "offline" and not args.no_cpu_baseline:      # a short sample of the CPU restatement on this box's host cores, same scene and resolution
                    cores = os.cpu_count() or 1
                    cv, nfr, dt = cpu_oracle_rate(ps2, cam2, min(args.cpu_seconds, 4.0), cores, w2, h2)
                    out["workloads"][-1]["cpu_baseline"] = {"value": cv, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": f"{nfr} x 1 spp x {w2}x{h2} frames in {dt:.1f} s"}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cv, nfr, dt = cpu_oracle_rate(ps, cam, args.cpu_seconds, cores, Wd, Hd)
            out["cpu_baseline"] = {"value": cv, "unit": "Msamples/s", "cores": cores, "kind": "port",
                                   "sample": f"{nfr} x 1 spp x {Wd}x{Hd} frames of the same scene in {dt:.1f} s (CPU restatement of the reference shaders, all host threads)"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
