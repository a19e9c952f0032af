"""Homography sharding across the GPUs of one box (replaces ``nn.DataParallel``: models/model_wrap.py:113-115,
export.py:265).

The N homographies of an image are independent until the two sums of ``combine_heatmap``
(export.py:61-62).  For a batch of B images every rank owns a block of the views of EVERY image (the blocks rotate
over the images so that uneven shares -- 100 views over 8 ranks = 4 x 13 + 4 x 12 -- even out), accumulates
``sum heat*mask`` and ``sum mask`` of its views into ``sums[B,2,H,W]`` and ONE collective per batch combines the
shards: a ``reduce_scatter`` over the image dimension when ``B % world == 0`` (rank r receives the summed planes of
the images ``[r*B/world, (r+1)*B/world)`` -- the only ones it divides and runs NMS / top-k on: half the bytes of an
all-reduce and no redundant work), otherwise an ``all_reduce`` followed by the same ownership rule.
Backend: on the GPUs of one box the exchange is ``PeerExchange`` -- the partial planes stay where they were computed,
the owner of an image adds them straight out of its peers' memory over NVLink (CUDA IPC mappings, ``csrc/peer.cu``)
inside the kernel that also divides, so no collective-library kernel runs on the data path; NCCL
(``exchange_sums``) remains for ranks on different hosts and as the cross-check.  The same functions run on ``gloo`` in
the CPU tests (tests/test_sharding_gloo.py drives ``plan_batch`` / ``exchange_sums`` / ``sharded_heatmap_sums`` -- the
code the product's ``HAExporter`` executes).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "balanced_view_plan", "owned_images", "owner_of_image", "plan_batch", "ShardPlan",
           "exchange_sums", "sharded_heatmap_sums", "world_info", "PeerExchange", "PeerUnavailable", "same_host"]


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
                         out: torch.Tensor = None, async_op: bool = False, exchange: Callable = None):
    """Run the sharded step: ``partial_fn(pass_plan, sums)`` adds (or, when ``not pass_plan.accumulate``, writes)
    this rank's partial sums of the pass's images into ``sums[b0:b0+nb]``; images without local views are zeroed
    here; then the exchange -- ``exchange(sums, plan)`` when given (the peer-memory path: ``PeerExchange``), else
    ``exchange_sums`` (NCCL / gloo).  Returns ``(owned, work, plan)`` like ``exchange_sums``."""
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
    if exchange is not None:
        owned, work = exchange(sums, plan)
    else:
        owned, work = exchange_sums(sums, plan, group, out=out, async_op=async_op)
    return owned, work, plan


# ------------------------------------------------------------------------------------------------------------
# the exchange over NVLink peer memory (one process per GPU of ONE box)
# ------------------------------------------------------------------------------------------------------------
def same_host(group=None) -> bool:
    """True when every rank of the group runs on this host (peer memory can be mapped through CUDA IPC)."""
    rank, world = world_info(group)
    if world == 1:
        return True
    import socket
    names = [None] * world
    dist.all_gather_object(names, socket.gethostname(), group=group)
    return len(set(names)) == 1


class PeerUnavailable(RuntimeError):
    """The ranks cannot map each other's memory (different hosts / containers, no P2P): use the NCCL exchange."""


class _DevMem:
    """A raw device allocation seen through ``__cuda_array_interface__`` (torch.as_tensor aliases it, no copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerExchange:
    """Two slots of partial planes ``sums[B,2,H,W]`` of this rank in a cudaMalloc'ed buffer every other rank of the box
    maps through CUDA IPC, plus the ``ready`` / ``consumed`` counters of the hand-shake (``csrc/peer.cu``).

    Per batch (``step`` = 0, 1, 2 ..., slot = step & 1), all calls only enqueue work:
      ``wait_reusable``  (main stream, before the slot is written) every owner has read step - 2's planes;
      ``publish``        (main stream, after the partial sums) tell every rank that this rank's planes are complete;
      ``finalize``       (any stream) wait for all ranks' planes of the slot, then ONE kernel loads the owned images'
                         planes from every peer, adds them in rank order and divides (export.py:61-63) into ``heat``;
                         finally tell every rank that this rank has consumed the slot.
    Construction is collective (all ranks, same arguments): it exchanges the IPC handles with one all-gather."""

    SLOTS, MAXP = 2, 16

    def __init__(self, num_images: int, H: int, W: int, device, group=None):
        import ctypes as C

        from . import _lib
        self._lib, self._C = _lib, C
        L = _lib.lib()
        self.rank, self.world = world_info(group)
        if self.world > self.MAXP:
            raise ValueError(f"PeerExchange supports up to {self.MAXP} ranks, got {self.world}")
        self.device, self.group = torch.device(device), group
        self.B, self.H, self.W = int(num_images), int(H), int(W)
        self.plane_bytes = self.B * 2 * self.H * self.W * 4
        self.flag_off = self.SLOTS * self.plane_bytes
        nbytes = self.flag_off + 2 * self.SLOTS * self.MAXP * 4
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.spb_peer_alloc(C.byref(ptr), nbytes), "peer_alloc")
            handle = (C.c_ubyte * 64)()
            _lib.check(L.spb_peer_export(ptr, handle), "peer_export")
            allh = [None] * self.world
            dist.all_gather_object(allh, bytes(handle), group=group)  # (any backend: NCCL on the box, gloo in tests)
            self.base = [0] * self.world
            self.base[self.rank] = int(ptr.value)
            self._own = int(ptr.value)
            err = None
            for p in range(self.world):
                if p == self.rank:
                    continue
                hb = (C.c_ubyte * 64)(*allh[p])
                q = C.c_void_p()
                try:
                    _lib.check(L.spb_peer_open(hb, C.byref(q)), f"peer_open (rank {p})")
                    self.base[p] = int(q.value)
                except RuntimeError as e:  # no peer access / IPC not permitted between these processes
                    err = e
                    break
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None, group=group)  # all ranks take the same decision
            if not all(oks):
                self._closed = False
                self.sums = []
                self.group = group
                self._release()
                raise PeerUnavailable(str(err) if err else "a peer rank could not map the shared buffers")
        self.sums = [torch.as_tensor(_DevMem(self._own + s * self.plane_bytes, (self.B, 2, self.H, self.W), "<f4"),
                                     device=self.device) for s in range(self.SLOTS)]
        self.step = 0
        self._closed = False
        dist.barrier(group=group)  # every mapping exists before anyone signals

    # flag (kind: 0 ready / 1 consumed) of `slot` that rank `writer` sets in rank `holder`'s memory
    def _flag(self, holder: int, kind: int, slot: int, writer: int) -> int:
        return self.base[holder] + self.flag_off + ((kind * self.SLOTS + slot) * self.MAXP + writer) * 4

    def _signal(self, kind: int, slot: int, value: int, stream: int):
        C = self._C
        arr = (C.c_void_p * self.world)(*[self._flag(q, kind, slot, self.rank) for q in range(self.world)])
        self._lib.check(self._lib.lib().spb_peer_signal(arr, self.world, value, stream), "peer_signal")

    def _wait(self, kind: int, slot: int, value: int, stream: int):
        self._lib.check(self._lib.lib().spb_peer_wait(self._flag(self.rank, kind, slot, 0), self.world, value, 0, stream),
                        "peer_wait")

    def wait_reusable(self, step: int, stream: int):
        if step >= self.SLOTS:
            self._wait(1, step % self.SLOTS, step - self.SLOTS + 1, stream)

    def publish(self, step: int, stream: int):
        self._signal(0, step % self.SLOTS, step + 1, stream)

    def finalize(self, step: int, lo: int, hi: int, heat: torch.Tensor, stream: int):
        C, slot = self._C, step % self.SLOTS
        if hi > lo:
            self._wait(0, slot, step + 1, stream)
            arr = (C.c_void_p * self.world)(*[self.base[p] + slot * self.plane_bytes for p in range(self.world)])
            self._lib.check(self._lib.lib().spb_ha_finalize_peers(arr, self.world, lo, hi - lo, heat.data_ptr(), self.H,
                                                                  self.W, stream), "ha_finalize_peers")
        self._signal(1, slot, step + 1, stream)

    def _release(self):
        L = self._lib.lib()
        for p, b in enumerate(self.base):
            if p != self.rank and b:
                L.spb_peer_close(b)
        self.base = [0] * self.world
        self.sums = []
        L.spb_peer_free(self._own)
        self._own = 0

    def close(self):
        """Collective: unmap the peers and free the buffer once every rank is done with it."""
        if self._closed:
            return
        self._closed = True
        try:
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier(group=self.group)
        except Exception:
            pass
        self._release()
