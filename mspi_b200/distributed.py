"""Clip sharding across GPUs (one process per GPU, torch.distributed).

Clips are independent in inference (eval-mode BatchNorm, model_utils.py:556-574 has no cross-sample op except the
batch mean inside loss_av), so the global batch is split contiguously by rank and the weights are replicated; there
is no data-path collective.  The only exchange is the gather of the [B_local, H, W] fp32 saliency maps (344 KB per
clip at 224x384) and the mean of the scalar loss_av — NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the global batch owned by `rank`; the first (global_batch % world) ranks get one
    extra clip, so any batch size (including fewer clips than ranks) is covered exactly once."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_shard(global_batch: int, world: int) -> int:
    return -(-global_batch // world)


class PendingGather:
    """Handle of an asynchronous gather_maps: .wait() makes the current stream wait for the collective and returns the maps."""

    def __init__(self, work, finish):
        self._work, self._finish = work, finish

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self._finish()


def gather_maps(local_maps: torch.Tensor, global_batch: int, group=None, async_op: bool = False):
    """All-gather the per-rank [b_local, H, W] maps into [global_batch, H, W] in global clip order.

    Ragged shards are padded to the largest shard for the collective (all_gather_into_tensor needs equal sizes)
    and the padding rows are dropped afterwards.  async_op=True returns a PendingGather: the collective runs on the
    communicator's own stream, so the next batch's kernels overlap it (wait() before the maps are read)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(global_batch, rank, world)
    if local_maps.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local_maps.shape[0]} maps, its shard is [{lo},{hi})")
    m = max_shard(global_batch, world)
    h, w = local_maps.shape[1:]
    send = local_maps
    if send.shape[0] != m:
        send = torch.zeros((m, h, w), dtype=local_maps.dtype, device=local_maps.device)
        send[: hi - lo] = local_maps
    recv = torch.empty((world * m, h, w), dtype=local_maps.dtype, device=local_maps.device)
    send = send.contiguous()
    work = dist.all_gather_into_tensor(recv, send, group=group, async_op=async_op)

    def finish(_keep=(send,)):
        if global_batch == world * m:
            return recv
        parts = []
        for r in range(world):
            a, b = shard_bounds(global_batch, r, world)
            parts.append(recv[r * m: r * m + (b - a)])
        return torch.cat(parts, 0)

    return PendingGather(work, finish) if async_op else finish()


def forward_sharded(forward: Callable, clips: torch.Tensor, audios: Optional[torch.Tensor], group=None):
    """Run `forward(clips_local, audios_local) -> (maps_local, loss_local)` on this rank's shard of a replicated
    global batch and return (all maps in global order, batch-mean loss).  The loss is the clip-weighted mean of the
    per-rank means, which equals the reference's mean over the global batch (model_utils.py:551)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = clips.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    if hi > lo:
        maps, loss = forward(clips[lo:hi], None if audios is None else audios[lo:hi])
        loss = torch.as_tensor(loss, dtype=torch.float32, device=maps.device).reshape(1) * (hi - lo)
    else:  # more ranks than clips: this rank contributes nothing
        h, w = clips.shape[-2:]
        maps = torch.empty((0, h, w), dtype=torch.float32, device=clips.device)
        loss = torch.zeros(1, dtype=torch.float32, device=clips.device)
    full = gather_maps(maps, n, group)
    dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return full, loss[0] / max(n, 1)


# ------------------------------------------------------------------------------------------ training (BASELINE config 5)
def allreduce_gradients(flat_grads: torch.Tensor, group=None) -> float:
    """The training step's only per-step collective: SUM all-reduce of the flat fp32 gradient buffer (411 tensors, 184 MB for
    MSPI-S3D) in one call — NCCL on GPUs, gloo in the CPU tests.  Returns the scale (1 / world_size) the optimiser kernel
    applies (`mspi_adamw_step(..., grad_scale)`), which together reproduce DistributedDataParallel's gradient averaging
    of per-rank batch-mean losses.  BatchNorm statistics stay per rank (the reference has no SyncBN)."""
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def broadcast_training_state(state, src: int = 0, group=None):
    """What DistributedDataParallel does when it wraps a module: every rank takes rank `src`'s parameters and buffers.  The
    decoder, SyncBlock and heads start from a random init (only the encoders are pretrained), so ranks that were not seeded
    identically would otherwise train different replicas on averaged gradients.  One broadcast for the flat fp32 parameter
    buffer, one per frozen tensor / BatchNorm buffer; the AdamW moments and the step count follow."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.broadcast(state.flat_p, src=src, group=group)
    dist.broadcast(state.flat_m, src=src, group=group)
    dist.broadcast(state.flat_v, src=src, group=group)
    for k in sorted(state.live):
        if k in state.offs:
            continue            # a view into flat_p
        t = state.live[k]
        if t.is_floating_point() or t.dtype in (torch.int64, torch.int32):
            dist.broadcast(t, src=src, group=group)
    step = torch.tensor([state.step_count], dtype=torch.int64, device=state.flat_p.device)
    dist.broadcast(step, src=src, group=group)
    state.step_count = int(step.item())


def parameters_checksum(state, group=None) -> bool:
    """True when every rank holds the same parameters (sum and sum of squares of the flat buffer agree bit for bit)."""
    s = torch.stack([state.flat_p.double().sum(), (state.flat_p.double() ** 2).sum()])
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return True
    lo, hi = s.clone(), s.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool(torch.equal(lo, hi))


def flat_layout(shapes, align: int = 4):
    """Offsets of tensors packed into one flat buffer with every tensor starting on an `align`-element boundary (16 bytes for
    fp32): (offsets, total).  The layout the training plan uses for parameters, gradients and both AdamW moments."""
    offs, n = [], 0
    for s in shapes:
        numel = 1
        for d in s:
            numel *= d
        offs.append(n)
        n += -(-numel // align) * align
    return offs, n
