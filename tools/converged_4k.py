#!/usr/bin/env python
"""Config 4: 3840x2160, 1024 spp converged render sharded over the GPUs of one box.

  python -m torch.distributed.run --nproc-per-node N tools/converged_4k.py [--spp 1024] [--workload helmet] [--check]

One process per GPU, scene replicated, the work plan of ohao_engine_b200/sharding.py (contiguous sample-index blocks,
sum-mode accumulation), ONE NCCL reduce of the RGBA32F image, resolve on rank 0.  Prints one JSON line: Msamples/s
(device time, max over ranks, reduce included).  --check also renders the whole job on rank 0 alone and reports the PSNR /
mean relative error of the sharded image against it (the same samples, only the fp32 summation order differs)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=1024); ap.add_argument("--width", type=int, default=3840); ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--workload", default="helmet", choices=["helmet", "cornell", "synthetic2m"]); ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch, torch.distributed as dist
    from bench import make_workload
    from ohao_engine_b200 import binding as B, sharding as S
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
        try: dist.init_process_group("nccl", device_id=torch.device("cuda", local)); dist.barrier()
        finally: os.dup2(saved, 1); os.close(saved)
    W, H = args.width, args.height
    ps, cam, desc = make_workload(args.workload)
    r = B.Renderer(W, H, device=local); r.set_scene(ps)
    v, p = cam.view(), cam.proj(W, H)
    items = S.plan(W, H, args.spp, world)[rank]
    ptr, _ = r.accum_dev_ptr()

    class _Alias:
        __cuda_array_interface__ = {"shape": (H * W * 4,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
    accum_t = torch.as_tensor(_Alias(), device=torch.device("cuda", local))
    S.render_plan(r, [S.WorkItem(it.tile, it.first_sample, min(it.nsamples, 4)) for it in items], v, p)      # warm-up (allocations, clocks)
    r.synchronize(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    r.reset_counters(); r.timer_start()
    S.render_plan(r, items, v, p)
    if world > 1:
        r.synchronize(); dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM); torch.cuda.synchronize()
    if rank == 0: r.resolve()
    ms = r.timer_stop()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    out = {"config": "4K converged", "workload": desc, "resolution": [W, H], "spp": args.spp, "n_gpus": world, "ms": ms,
           "value": W * H * args.spp / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "plan": f"{world} sample-index blocks, sum mode, one NCCL reduce of {W * H * 16 / 1e6:.0f} MB"}
    if args.check and rank == 0:
        acc, _, _ = r.readback_hdr_buffers(want_aov=False)
        mean_sh = acc[..., :3] / np.maximum(acc[..., 3:4], 1.0)
        one = B.Renderer(W, H, device=local); one.set_scene(ps)
        done = 0
        while done < args.spp:
            k = min(64, args.spp - done); one.render(v, p, k); done += k
        ref, _, _ = one.readback_hdr_buffers(want_aov=False)
        ref = ref[..., :3].astype(np.float64); got = mean_sh.astype(np.float64)
        mse = float(np.mean((np.clip(ref, 0, 1) - np.clip(got, 0, 1)) ** 2))
        out["check"] = {"psnr_db_vs_single_gpu": 99.0 if mse == 0 else float(10 * np.log10(1.0 / mse)),
                        "mean_rel_err": float(np.abs(ref - got).mean() / max(ref.mean(), 1e-9))}
    if rank == 0: print(json.dumps(out), flush=True)
    if world > 1: dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
