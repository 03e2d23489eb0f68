# -*- coding: utf-8 -*-
"""
Multi-GPU plumbing: resamples are independent, so every rank (one process per
GPU) runs a contiguous block of resample ids against its own replica of X / Y
and the results are combined once at the end -- an all-gather of the
per-resample outputs (permuted singular values, bootstrap distributions) and
an all-reduce of the bootstrap accumulators (the parent-side sum of
pyls/base.py:510-511).  Works on any ``torch.distributed`` backend (NCCL on
the GPUs, gloo in the CPU tests).
"""

import torch
import torch.distributed as dist

from .resample import shard_range


def world():
    """(rank, world_size) of the default process group, (0, 1) without one."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def my_block(n):
    """(first, count) of the resample ids this rank owns."""
    rank, size = world()
    return shard_range(n, rank, size)


def gather_resamples(local, n_total):
    """
    All-gathers per-resample results along axis 0.

    ``local`` holds this rank's block (count_r, ...) of the ``n_total``
    resamples split by :func:`shard_range`; the result is the full
    (n_total, ...) tensor in resample-id order on every rank.
    """
    rank, size = world()
    if size == 1:
        return local
    counts = [shard_range(n_total, r, size)[1] for r in range(size)]
    width = max(counts)
    tail = local.shape[1:]
    padded = local.new_zeros((width,) + tuple(tail))
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(size)]
    dist.all_gather(parts, padded.contiguous())
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def reduce_sum(*tensors):
    """In-place all-reduce (sum) of every tensor given."""
    _, size = world()
    if size > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors
