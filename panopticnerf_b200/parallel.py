"""Ray sharding across the GPUs of one box.

Rays are independent, so the render path shards with no data-path collective: rank r renders the
contiguous ray range [r*ceil(R/G), ...) with the same kernels, and ONE all-gather of the per-ray output
maps (the reference has no collective on this path; BASELINE.json's north_star adds this one) rebuilds
the full image on every rank.  Results are bit-identical to the single-GPU render because every kernel
is deterministic and per-ray.

torch.distributed is only plumbing here (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Optional, Tuple

import torch
import torch.distributed as dist

DEFAULT_KEYS = ("rgb_map", "depth_map", "acc_map")


def shard_range(R: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous equal ranges of ceil(R/world) rays; the last ranks may be short or empty."""
    per = (R + world - 1) // world
    lo = min(rank * per, R)
    return lo, min(lo + per, R)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Per-ray tensors (leading dim R) are sliced; scene tensors (boxes, aabb) are replicated."""
    R = batch["rays"].shape[0]
    lo, hi = shard_range(R, rank, world)
    out = {}
    for k, v in batch.items():
        per_ray = torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == R and k in (
            "rays", "near", "far", "u", "u_fine")
        out[k] = v[lo:hi].contiguous() if per_ray else v
    return out


def all_gather_maps(local: Dict[str, torch.Tensor], R: int, keys: Iterable[str] = DEFAULT_KEYS,
                    group: Optional[dist.ProcessGroup] = None) -> Dict[str, torch.Tensor]:
    """One all_gather_into_tensor of the selected per-ray maps, packed as a single [per, F] fp32 tile
    per rank (padded to ceil(R/world) rows), then unpacked to full-image tensors on every rank."""
    world = dist.get_world_size(group)
    per = (R + world - 1) // world
    keys = [k for k in keys if k in local]
    cols = [local[k].reshape(local[k].shape[0], -1).to(torch.float32) for k in keys]
    widths = [c.shape[1] for c in cols]
    n_loc = cols[0].shape[0]
    tile = torch.zeros(per, sum(widths), dtype=torch.float32, device=cols[0].device)
    if n_loc:
        tile[:n_loc] = torch.cat(cols, 1)
    full = torch.empty(world * per, sum(widths), dtype=torch.float32, device=tile.device)
    dist.all_gather_into_tensor(full, tile, group=group)
    full = full[:R]
    out, c0 = {}, 0
    for k, w in zip(keys, widths):
        shape = (R,) + tuple(local[k].shape[1:])
        out[k] = full[:, c0:c0 + w].reshape(shape).to(local[k].dtype)
        c0 += w
    return out


def render_sharded(render_fn: Callable[[Dict[str, torch.Tensor]], Dict[str, torch.Tensor]],
                   batch: Dict[str, torch.Tensor], keys: Iterable[str] = DEFAULT_KEYS,
                   group: Optional[dist.ProcessGroup] = None) -> Dict[str, torch.Tensor]:
    """Render this rank's ray shard with `render_fn` (Renderer.render) and all-gather the maps."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    R = batch["rays"].shape[0]
    local = render_fn(shard_batch(batch, rank, world))
    return all_gather_maps(local, R, keys, group)
