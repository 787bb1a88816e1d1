"""Multi-GPU sharding of the path-tracing hot path: one process per GPU, scene replicated.

Work item = (tile rectangle, contiguous sample-index block).  Every rank accumulates a SUM image
(``ohb_set_accum_mode(ctx, 1)``: rgb = sum of radiance, a = sample count) so partial images add exactly;
ONE reduce of the RGBA32F image closes a render (NCCL over NVLink on GPUs, gloo in the CPU tests).  The sample
sequence is a pure function of (pixel, sampleIndex) (sampler_sobol.glsl:60-82), so any partition of
pixels x sampleIndices renders the same set of samples as one GPU; only the fp32 summation order differs.
Inverse-fit probes (finite-difference renders) are independent jobs distributed round-robin.
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence, Tuple


@dataclasses.dataclass(frozen=True)
class WorkItem:
    tile: Tuple[int, int, int, int]      # x0, y0, w, h
    first_sample: int                    # absolute sample index of the block's first sample
    nsamples: int


def sample_blocks(total_spp: int, parts: int, seed: int = 0) -> List[Tuple[int, int]]:
    """Split sample indices [seed, seed+total_spp) into `parts` contiguous blocks (first, count); sizes differ by <= 1."""
    base, extra = divmod(total_spp, parts)
    out, first = [], seed
    for g in range(parts):
        n = base + (1 if g < extra else 0)
        out.append((first, n)); first += n
    return out


def tiles(width: int, height: int, tile: int = 256) -> List[Tuple[int, int, int, int]]:
    """Row-major tile rectangles covering the frame (ragged at the right / bottom edges)."""
    return [(x, y, min(tile, width - x), min(tile, height - y)) for y in range(0, height, tile) for x in range(0, width, tile)]


def plan(width: int, height: int, total_spp: int, world: int, seed: int = 0, tile: int = 256, spp_parts: int | None = None) -> List[List[WorkItem]]:
    """Work list per rank.  `spp_parts` ranks-groups split the sample range; inside a group tiles go round-robin.
    Default: split samples when there are at least as many samples as ranks (no tile seams, best balance),
    otherwise split tiles only."""
    if spp_parts is None:
        spp_parts = world if total_spp >= world else 1
    if world % spp_parts:
        raise ValueError("spp_parts must divide the world size")
    tile_parts = world // spp_parts
    blocks = sample_blocks(total_spp, spp_parts, seed)
    ts = tiles(width, height, tile) if tile_parts > 1 else [(0, 0, width, height)]
    out: List[List[WorkItem]] = [[] for _ in range(world)]
    for r in range(world):
        sp, tp = divmod(r, tile_parts)
        first, n = blocks[sp]
        if n == 0:
            continue
        for i, t in enumerate(ts):
            if i % tile_parts == tp:
                out[r].append(WorkItem(t, first, n))
    return out


def jobs_for_rank(njobs: int, world: int, rank: int) -> List[int]:
    """Round-robin distribution of independent jobs (inverse-fit probes)."""
    return list(range(rank, njobs, world))


def render_plan(renderer, items: Sequence[WorkItem], view, proj, max_batch: int = 64):
    """Run a rank's work list on a ``binding.Renderer`` in sum mode."""
    renderer.set_accum_mode(True)
    renderer.clear_accum()          # pixels outside this rank's tiles must be zero when the images are summed
    for it in items:
        renderer.set_tile(*it.tile)
        renderer.set_render_seed(it.first_sample)          # also resets the per-context history counter ...
        done = 0
        while done < it.nsamples:
            k = min(max_batch, it.nsamples - done)
            renderer.render(view, proj, k); done += k
    renderer.set_tile(0, 0, renderer.width, renderer.height)


def reduce_sum_image(accum, dst: int = 0):
    """ONE collective per image: sum-reduce the (H, W, 4) float32 accumulation tensor onto rank `dst`."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


def gather_scalars(values: Sequence[float], njobs: int, world: int, rank: int):
    """All ranks learn every job's scalar result (one double per probe): returns a list of length njobs."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(njobs, dtype=torch.float64)
    for j, v in zip(jobs_for_rank(njobs, world, rank), values):
        buf[j] = v
    if dist.is_initialized() and world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf.tolist()
