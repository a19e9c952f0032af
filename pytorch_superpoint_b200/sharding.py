"""Homography sharding across the GPUs of one box (replaces ``nn.DataParallel``: models/model_wrap.py:113-115,
export.py:265).

The N homographies of an image are independent until the two sums of ``combine_heatmap``
(export.py:61-62).  Rank r owns a contiguous block of view indices (rank 0's block holds the identity
view i = 0), accumulates ``sum heat*mask`` and ``sum mask`` for its block into ``sums[B,2,H,W]`` and ONE
``all_reduce(SUM)`` of that buffer per batch of B images (0.61 MB per image at 240x320) combines the shards;
the division, NMS and top-k of image b then run on rank ``b % world`` only.  Backend: NCCL on GPUs
(NVLink 5 / NVSwitch); the same code runs on ``gloo`` for the CPU tests.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "balanced_view_plan", "owner_of_image", "sharded_heatmap_sums", "world_info"]


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(num_views: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of view indices for ``rank``; sizes differ by at most one
    (100 views over 8 ranks = 4 x 13 + 4 x 12)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(num_views, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_view_plan(num_images: int, num_views: int, rank: int, world: int):
    """Homography sharding balanced over a BATCH: image b gives rank r the block ``shard_bounds(N, (r + b) % world)``,
    so the larger shares (13 of 100 views over 8 ranks, against 12) rotate over the images and every rank ends up with
    the same number of views per step when ``num_images % world == 0`` (plain per-image blocks leave the 13-view ranks
    4 % more work than the average, and the step runs at their pace).  Every (image, view) pair still belongs to
    exactly one rank and every image is spread over all ranks.

    Returns ``(select, view_image, view_offsets, max_views)``: ``select`` indexes the image-major flat list of
    ``num_images * num_views`` (image, view) pairs, ``view_image[v]`` is the image of the v-th selected view,
    image b owns ``[view_offsets[b], view_offsets[b+1])`` (int64 / int32 / int32 tensors on the CPU)."""
    select, view_image, offsets = [], [], [0]
    most = 0
    for b in range(num_images):
        lo, hi = shard_bounds(num_views, (rank + b) % world, world)
        select.extend(range(b * num_views + lo, b * num_views + hi))
        view_image.extend([b] * (hi - lo))
        offsets.append(offsets[-1] + hi - lo)
        most = max(most, hi - lo)
    return (torch.tensor(select, dtype=torch.int64), torch.tensor(view_image, dtype=torch.int32),
            torch.tensor(offsets, dtype=torch.int32), most)


def owner_of_image(b: int, world: int) -> int:
    return b % world


def sharded_heatmap_sums(partial_fn: Callable[[int, int, int], torch.Tensor], num_images: int, num_views: int,
                         group=None) -> torch.Tensor:
    """``partial_fn(b, lo, hi)`` returns this rank's ``[2,H,W]`` sums of image b over views [lo, hi)
    (a zero tensor if the block is empty).  Returns the all-reduced ``[B,2,H,W]`` buffer."""
    rank, world = world_info(group)
    lo, hi = shard_bounds(num_views, rank, world)
    parts = [partial_fn(b, lo, hi) for b in range(num_images)]
    sums = torch.stack(parts, dim=0) if len(parts) > 1 else parts[0][None]
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums
