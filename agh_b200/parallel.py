"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPU, gloo in
the CPU tests).  The rollout path shards instances in contiguous blocks (mirrors eval.py:45-46) with
no data-path collective; the only exchanges are
  C2: an all-gather of the per-instance per-fleet costs [B/G,10] at the end of a rollout, and
  C1: one flat-bucket all-reduce(avg) of the 705,408 f32 gradients per training step,
which replace the reference's vestigial nn.DataParallel / mp.Pool paths (run.py:78-79, eval.py:58-67).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of rank `rank`: the first n % G ranks take one extra item."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Gather row blocks of unequal length from all ranks into [n_total, ...] (every rank gets the result)."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_bounds(n_total, r, ws) for r in range(ws)]
    longest = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((longest,) + tuple(local.shape[1:]))
    pad[:local.size(0)] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    return torch.cat([bufs[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)


def allreduce_grads_(params, buffers=()) -> None:
    """C1: average gradients across ranks through ONE flat f32 bucket (2.82 MB for this model).

    `buffers`: floating-point module buffers (the BatchNorm running statistics) averaged in the same bucket, so that
    every rank keeps identical weights AND identical eval-mode BatchNorm folds: train-mode BatchNorm uses per-rank batch
    statistics (the reference's DataParallel semantics, graph_encoder.py:137-138), which would otherwise let the
    running_mean / running_var of the ranks drift apart after the first step."""
    rank, ws = world()
    if ws == 1:
        return
    tensors = [p.grad for p in params if p.grad is not None] + [b for b in buffers if b.is_floating_point()]
    if not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(ws)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
