"""Multi-GPU layer: independent (batch item, view) renders are sharded across ranks, one process per GPU.

The reference scales only by Lightning DDP with batch 1 per GPU (/root/reference/src/main.py:117-130,
README.md:133) and renders the views of a batch item one after another in Python
(/root/reference/src/model/decoder/decoder_splatting_cuda.py:44-59).  Each view render is independent
given the Gaussians, so the path shards with NO data-path collective; the only collective is the
all-reduce of the scalar loss / metrics (NCCL on GPUs, gloo in the CPU tests).  If one scene's views are
split across ranks AND Gaussian gradients are needed, ``all_reduce_gradients`` sums the per-rank
gradient blocks (85 floats per Gaussian).
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_views(num_views: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """Round-robin assignment of view indices to this rank: rank r renders views
    r, r + world, r + 2 world, ...  Every view is rendered by exactly one rank."""
    if rank is None or world_size is None:
        rank, world_size = world()
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, num_views, world_size))


def shard_view_groups(num_views: int, group: int, rank: Optional[int] = None,
                      world_size: Optional[int] = None) -> List[int]:
    """Like ``shard_views`` but whole groups of ``group`` consecutive views stay on one rank -- the six cube faces of a
    panorama (/root/reference/src/model/model_wrapper_erp.py:336-345 renders targets x 6 faces as one view list), so
    that each rank can render its panoramas in batched passes that share the camera centre (``rasterize_views``).
    Groups are dealt round-robin; a trailing partial group is a group of its own."""
    if group <= 0:
        raise ValueError("group must be positive")
    if rank is None or world_size is None:
        rank, world_size = world()
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    n_groups = (num_views + group - 1) // group
    out: List[int] = []
    for g in range(rank, n_groups, world_size):
        out.extend(range(g * group, min((g + 1) * group, num_views)))
    return out


def all_reduce_loss(loss: Tensor, average: bool = False) -> Tensor:
    """Sum (or mean) of a scalar loss over ranks -- the only collective on the hot path."""
    _, ws = world()
    if ws == 1:
        return loss
    out = loss.detach().clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out / ws if average else out


class AsyncLossReducer:
    """The loss all-reduce off the critical path: the reduction of step i is issued asynchronously (NCCL runs it on
    its own stream, ordered after the loss of step i) and is only waited for when its ring slot is reused ``depth``
    steps later or when the value is read, so rendering of step i+1 never waits for the slowest rank's step i."""

    def __init__(self, device, depth: int = 2) -> None:
        self.bufs = [torch.zeros(1, dtype=torch.float32, device=device) for _ in range(depth)]
        self.works = [None] * depth
        self.n = 0

    def submit(self, loss: Tensor) -> None:
        j = self.n % len(self.bufs)
        if self.works[j] is not None:
            self.works[j].wait()
            self.works[j] = None
        self.bufs[j].copy_(loss.detach().reshape(1))
        _, ws = world()
        if ws > 1:
            self.works[j] = dist.all_reduce(self.bufs[j], op=dist.ReduceOp.SUM, async_op=True)
        self.n += 1

    def latest(self) -> Tensor:
        """Summed loss of the most recently submitted step (waits for its all-reduce)."""
        if self.n == 0:
            raise RuntimeError("no loss submitted yet")
        j = (self.n - 1) % len(self.bufs)
        if self.works[j] is not None:
            self.works[j].wait()
            self.works[j] = None
        return self.bufs[j]

    def flush(self) -> None:
        for j, w in enumerate(self.works):
            if w is not None:
                w.wait()
                self.works[j] = None


def all_reduce_gradients(grads: Iterable[Optional[Tensor]]) -> None:
    """In-place sum over ranks of per-Gaussian gradient tensors (only needed when the views of ONE scene are
    split across ranks and the Gaussians are replicated)."""
    _, ws = world()
    if ws == 1:
        return
    for g in grads:
        if g is not None:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)


def broadcast_scene(tensors: Sequence[Tensor], src: int = 0) -> None:
    """Replicate one scene's Gaussian tensors (means, covariances, opacities, harmonics: 340-352 B per Gaussian) from
    rank ``src`` to every rank, in place.  For the replicated-scene configurations (the video path renders many frames
    of ONE scene, /root/reference/src/model/model_wrapper_erp.py:412-432) this replaces one PCIe upload per GPU by one
    upload plus a broadcast over NVLink / NVSwitch (NCCL; gloo in the CPU tests)."""
    _, ws = world()
    if ws == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src)


def upload_and_broadcast_scene(host: Optional[Sequence[Tensor]], bufs: Sequence[Tensor], src: int = 0,
                               chunk_bytes: int = 64 << 20, copy_stream: Optional["torch.cuda.Stream"] = None) -> None:
    """Replicated-scene upload: rank ``src`` holds the scene in (pinned) host tensors ``host``; every rank ends up with it in
    its device buffers ``bufs`` (same shapes).  The tensors are cut into chunks of ``chunk_bytes``: rank ``src`` uploads chunk
    k+1 over PCIe on a copy stream while chunk k is being broadcast over NVLink / NVSwitch (NCCL), so the broadcast costs
    (almost) nothing on top of the single upload.  Stream-ordered on the current stream; other ranks pass ``host=None``."""
    rank, ws = world()
    cur = torch.cuda.current_stream(bufs[0].device)
    if rank == src:
        if host is None:
            raise ValueError("the source rank needs the host tensors")
        side = copy_stream if copy_stream is not None else torch.cuda.Stream(device=bufs[0].device)
        side.wait_stream(cur)   # the buffers may still be read by earlier work on the current stream
    for i, b in enumerate(bufs):
        flat = b.view(-1)
        per = max(1, chunk_bytes // flat.element_size())
        hflat = host[i].view(-1) if rank == src else None
        for o in range(0, flat.numel(), per):
            piece = flat[o:o + per]
            if rank == src:
                with torch.cuda.stream(side):
                    piece.copy_(hflat[o:o + per], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(side)
                cur.wait_event(ev)
            if ws > 1:
                dist.broadcast(piece, src=src)


def gather_views(local: Tensor, num_views: int) -> Tensor:
    """Reassemble the [num_views, ...] stack from the round-robin shards ``local`` [len(shard), ...] held by each
    rank (evaluation / video paths that need every frame on every rank)."""
    rank, ws = world()
    if ws == 1:
        return local
    per = (num_views + ws - 1) // ws
    pad = torch.zeros((per, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    out = torch.empty((num_views, *local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(ws):
        idx = shard_views(num_views, r, ws)
        out[idx] = parts[r][: len(idx)]
    return out


def render_views_sharded(render_one: Callable[[int], Tensor], num_views: int) -> Tuple[List[int], List[Tensor]]:
    """Run ``render_one(view_index)`` for this rank's share of the views.  Returns (indices, images)."""
    idx = shard_views(num_views)
    return idx, [render_one(i) for i in idx]
