"""Homography sharding across the GPUs of one box (replaces ``nn.DataParallel``: models/model_wrap.py:113-115,
export.py:265).

The N homographies of an image are independent until the two sums of ``combine_heatmap``
(export.py:61-62).  For a batch of B images every rank owns a block of the views of EVERY image (the blocks rotate
over the images so that uneven shares -- 100 views over 8 ranks = 4 x 13 + 4 x 12 -- even out), accumulates
``sum heat*mask`` and ``sum mask`` of its views into ``sums[B,2,H,W]`` and ONE collective per batch combines the
shards: a ``reduce_scatter`` over the image dimension when ``B % world == 0`` (rank r receives the summed planes of
the images ``[r*B/world, (r+1)*B/world)`` -- the only ones it divides and runs NMS / top-k on: half the bytes of an
all-reduce and no redundant work), otherwise an ``all_reduce`` followed by the same ownership rule.
Backend: NCCL on GPUs (NVLink 5 / NVSwitch); the same functions run on ``gloo`` in the CPU tests
(tests/test_sharding_gloo.py drives ``plan_batch`` / ``exchange_sums`` / ``sharded_heatmap_sums`` -- the code the
product's ``HAExporter`` executes).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "balanced_view_plan", "owned_images", "owner_of_image", "plan_batch", "ShardPlan",
           "exchange_sums", "sharded_heatmap_sums", "world_info"]


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(num_views: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``num_views`` items for ``rank``; sizes differ by at most one
    (100 views over 8 ranks = 4 x 13 + 4 x 12)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(num_views, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owned_images(num_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Images [lo, hi) whose division / NMS / top-k run on ``rank``: contiguous blocks, i.e. exactly the chunk a
    ``reduce_scatter`` over the image dimension delivers when ``num_images % world == 0``."""
    return shard_bounds(num_images, rank, world)


def owner_of_image(b: int, world: int, num_images: int = 0) -> int:
    """Rank that finalises image b of a batch of ``num_images``."""
    if num_images <= 0:
        raise ValueError("owner_of_image needs the batch size (contiguous ownership blocks)")
    for r in range(world):
        lo, hi = shard_bounds(num_images, r, world)
        if lo <= b < hi:
            return r
    raise ValueError(f"image {b} outside batch of {num_images}")


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


@dataclass
class PassPlan:
    """One set of kernel launches: the images ``[b0, b0 + nb)`` of the batch with ``view_offsets[nb+1]`` views each
    (``view_image`` relative to b0, ``view_mat`` = index of every view's matrix in the image-major ``[B*N]`` list,
    or ``view % N`` into a shared ``[N]`` set), ``accumulate`` = add to the planes an earlier pass of the same image
    wrote (an image whose views do not fit one pass)."""
    b0: int
    nb: int
    view_image: torch.Tensor
    view_mat: torch.Tensor
    view_offsets: torch.Tensor
    max_views: int
    accumulate: bool

    @property
    def num_views(self) -> int:
        return int(self.view_image.numel())


@dataclass
class ShardPlan:
    """What THIS rank computes and finalises for a batch of B images x N views."""
    num_images: int
    num_views: int
    rank: int
    world: int
    ranges: List[Tuple[int, int]]          # per image: this rank's view block [lo, hi)
    own: Tuple[int, int]                   # images this rank divides / extracts key-points from
    scatter: bool                          # exchange = reduce_scatter (B % world == 0) instead of all_reduce
    passes: List[PassPlan] = field(default_factory=list)

    @property
    def local_views(self) -> int:
        return sum(hi - lo for lo, hi in self.ranges)


def plan_batch(num_images: int, num_views: int, rank: int, world: int, max_pass_views: int = 0) -> ShardPlan:
    """Rotated view blocks for every image of the batch (see ``balanced_view_plan``), cut into passes of at most
    ``max_pass_views`` views (0 = one pass): whole images while they fit, an image with more views than a pass holds
    is spread over several accumulating passes."""
    B, N = int(num_images), int(num_views)
    ranges = [shard_bounds(N, (rank + b) % world, world) for b in range(B)]
    plan = ShardPlan(B, N, rank, world, ranges, owned_images(B, rank, world), scatter=world > 1 and B % world == 0)
    cap = max_pass_views if max_pass_views and max_pass_views > 0 else 1 << 30
    cur: List[Tuple[int, int, int]] = []  # (image, lo, hi) of the pass being filled: consecutive images
    used = 0

    def close(accumulate=False):
        nonlocal used
        if cur and any(hi > lo for _, lo, hi in cur):
            b0 = cur[0][0]
            vi, vm, vo = [], [], [0]
            for b, lo, hi in cur:
                vi.extend([b - b0] * (hi - lo))
                vm.extend(range(b * N + lo, b * N + hi))
                vo.append(vo[-1] + hi - lo)
            plan.passes.append(PassPlan(b0, len(cur), torch.tensor(vi, dtype=torch.int32),
                                        torch.tensor(vm, dtype=torch.int32), torch.tensor(vo, dtype=torch.int32),
                                        max(hi - lo for _, lo, hi in cur), accumulate))
        cur.clear()
        used = 0

    for b, (lo, hi) in enumerate(ranges):
        n = hi - lo
        if n > cap:  # more views than a pass holds: the image travels alone, in accumulating slices
            close()
            first = True
            while lo < hi:
                m = min(cap, hi - lo)
                cur.append((b, lo, lo + m))
                close(accumulate=not first)
                lo, first = lo + m, False
            continue
        if cur and used + n > cap:
            close()
        cur.append((b, lo, hi))  # (an image without local views rides along: its planes come out as zeros)
        used += n
    close()
    return plan


def exchange_sums(sums: torch.Tensor, plan: ShardPlan, group=None, out: torch.Tensor = None,
                  async_op: bool = False):
    """The ONE collective of the path.  ``sums`` [B,2,H,W] = this rank's partial planes of every image of the batch.
    Returns ``(owned, work)``: ``owned`` [own_hi - own_lo, 2, H, W] are the fully summed planes of the images this
    rank finalises (a view of ``sums`` after an all-reduce, ``out`` after a reduce-scatter); ``work`` is the async
    handle (None when synchronous or single-rank)."""
    lo, hi = plan.own
    if plan.world == 1:
        return sums, None
    if plan.scatter:
        if out is None:
            out = torch.empty((hi - lo,) + tuple(sums.shape[1:]), dtype=sums.dtype, device=sums.device)
        work = dist.reduce_scatter_tensor(out, sums, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return out, work
    work = dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return sums[lo:hi], work


def sharded_heatmap_sums(partial_fn: Callable[[PassPlan, torch.Tensor], None], num_images: int, num_views: int,
                         sums: torch.Tensor, group=None, max_pass_views: int = 0, plan: ShardPlan = None,
                         out: torch.Tensor = None, async_op: bool = False):
    """Run the sharded step: ``partial_fn(pass_plan, sums)`` adds (or, when ``not pass_plan.accumulate``, writes)
    this rank's partial sums of the pass's images into ``sums[b0:b0+nb]``; images without local views are zeroed
    here; then the exchange.  Returns ``(owned, work, plan)`` like ``exchange_sums``."""
    rank, world = world_info(group)
    if plan is None:
        plan = plan_batch(num_images, num_views, rank, world, max_pass_views)
    touched = set()
    for p in plan.passes:
        partial_fn(p, sums)
        touched.update(range(p.b0, p.b0 + p.nb))
    for b in range(num_images):
        if b not in touched:  # fewer views than ranks: this rank has none of image b
            sums[b].zero_()
    owned, work = exchange_sums(sums, plan, group, out=out, async_op=async_op)
    return owned, work, plan
