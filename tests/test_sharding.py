"""Multi-rank logic on CPU: partition properties + a world_size-2 gloo run whose per-rank partial SUM images
(rendered by the host emulator of the kernels, test tooling) reduce to the single-rank render."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ohao_engine_b200 import sharding as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_blocks_and_plan_partition_the_work_exactly():
    for spp, parts, seed in [(64, 8, 0), (10, 4, 5), (3, 8, 0), (1024, 8, 7)]:
        b = S.sample_blocks(spp, parts, seed)
        assert sum(n for _, n in b) == spp and b[0][0] == seed
        assert all(b[i][0] + b[i][1] == b[i + 1][0] for i in range(parts - 1))
        assert max(n for _, n in b) - min(n for _, n in b) <= 1
    for (W, H, spp, world, parts) in [(1920, 1080, 1024, 8, None), (3840, 2160, 4, 8, None), (640, 360, 16, 4, 2), (100, 70, 1, 2, 1)]:
        cover = np.zeros((H, W), np.int64)
        for items in S.plan(W, H, spp, world, spp_parts=parts, tile=64):
            for it in items:
                x, y, w, h = it.tile
                cover[y:y + h, x:x + w] += it.nsamples
        assert (cover == spp).all()
    assert sorted(sum((S.jobs_for_rank(72, 8, r) for r in range(8)), [])) == list(range(72))
    with pytest.raises(ValueError):
        S.plan(64, 64, 16, 8, spp_parts=3)


def _worker(rank, world, port, W, H, spp, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ohao_engine_b200 import scenes
    from tests.emul import emul_py as E
    ps, cam = scenes.cornell_box(), scenes.cornell_camera()
    esc = E.EmulScene(ps)
    acc = np.zeros((H, W, 4), np.float32)
    for it in S.plan(W, H, spp, world, spp_parts=1 if rank < 0 else None)[rank]:      # spp split across the 2 ranks
        part = esc.render_offline(cam.view(), cam.proj(W, H), W, H, it.nsamples, first_sample=it.first_sample, tile=it.tile)["accum"]
        x, y, w, h = it.tile
        acc[y:y + h, x:x + w, :3] += part[y:y + h, x:x + w, :3] * part[y:y + h, x:x + w, 3:4]     # running mean -> sum
        acc[y:y + h, x:x + w, 3] += part[y:y + h, x:x + w, 3]
    t = torch.from_numpy(acc)
    S.reduce_sum_image(t, dst=0)
    losses = S.gather_scalars([float(j * j) for j in S.jobs_for_rank(7, world, rank)], 7, world, rank)
    assert losses == [float(j * j) for j in range(7)]
    if rank == 0:
        np.save(out, t.numpy())
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_gloo_reduce_matches_single_rank(tmp_path):
    W, H, spp = 48, 27, 4
    out = str(tmp_path / "acc.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, W, H, spp, out), nprocs=2, join=True)
    acc = np.load(out)
    from ohao_engine_b200 import scenes
    from tests.emul import emul_py as E
    ps, cam = scenes.cornell_box(), scenes.cornell_camera()
    full = E.EmulScene(ps).render_offline(cam.view(), cam.proj(W, H), W, H, spp)["accum"]
    assert (acc[..., 3] == spp).all()
    assert np.allclose(acc[..., :3] / spp, full[..., :3], rtol=1e-5, atol=1e-6)
