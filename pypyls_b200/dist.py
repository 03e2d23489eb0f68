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


def broadcast_from_root(eng, a, shape):
    """Device copy of ``a`` (a host or device array of ``shape``) taken from rank 0:
    rank 0 uploads its array, the other ranks receive it over NCCL / NVLink instead of
    pushing their own replicas through the host's memory and PCIe links at the same time."""
    rank, size = world()
    if size == 1:
        return eng.to_device(a)
    if rank == 0:
        t = eng.to_device(a)
        if tuple(t.shape) != tuple(shape):
            raise ValueError('expected an array of shape {}'.format(tuple(shape)))
    else:
        t = torch.empty(tuple(shape), dtype=torch.float64, device=eng.device)
    dist.broadcast(t, src=0)
    return t


def my_block(n):
    """(first, count) of the resample ids this rank owns."""
    rank, size = world()
    return shard_range(n, rank, size)


def gather_resamples(local, n_total):
    """
    All-gathers per-resample results along axis 0.

    ``local`` holds this rank's block (count_r, ...) of the ``n_total``
    resamples split by :func:`shard_range`; the result is the full
    (n_total, ...) tensor in resample-id order on every rank.  Equal blocks
    (the usual case) land directly in the preallocated result
    (``all_gather_into_tensor``: no padding, no concatenation).
    """
    rank, size = world()
    if size == 1:
        return local
    counts = [shard_range(n_total, r, size)[1] for r in range(size)]
    tail = tuple(local.shape[1:])
    local = local.contiguous()
    if min(counts) == max(counts):
        out = local.new_empty((n_total,) + tail)
        dist.all_gather_into_tensor(out, local)
        return out
    width = max(counts)
    padded = local.new_zeros((width,) + tail)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(size)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def gather_to_root(local, n_total, root=0):
    """Rank `root` receives the full (n_total, ...) tensor in resample-id
    order, the other ranks None: what the host of rank `root` needs of the
    per-resample outputs, moved once instead of to every rank."""
    rank, size = world()
    if size == 1:
        return local
    counts = [shard_range(n_total, r, size)[1] for r in range(size)]
    tail = tuple(local.shape[1:])
    local = local.contiguous()
    even = min(counts) == max(counts)
    if rank != root:
        if even:
            dist.gather(local, None, dst=root)
        else:
            dist.send(local, dst=root)
        return None
    out = local.new_empty((n_total,) + tail)
    offs = [shard_range(n_total, r, size)[0] for r in range(size)]
    parts = [out[o:o + c] for o, c in zip(offs, counts)]
    if even:
        dist.gather(local, parts, dst=root)
    else:
        for r in range(size):
            if r == root:
                parts[r].copy_(local)
            else:
                dist.recv(parts[r], src=r)
    return out


def exchange_series(local_t, n_total):
    """
    Deals the series of a per-resample output over the ranks.

    ``local_t`` is this rank's block TRANSPOSED to series-major,
    (n_series, count_r).  Rank q receives series
    ``shard_range(n_series, q, size)`` of EVERY rank's block and returns them
    as one (my_series, n_total) tensor (resample-id order along axis 1) -- one
    all-to-all in which every rank sends and receives about 1/size of the whole
    output, instead of an all-gather that hands all of it to everyone.
    """
    rank, size = world()
    if size == 1:
        return local_t
    n_series, mine = int(local_t.shape[0]), int(local_t.shape[1])
    counts = [shard_range(n_total, r, size)[1] for r in range(size)]
    offs = [shard_range(n_total, r, size)[0] for r in range(size)]
    sblocks = [shard_range(n_series, r, size) for r in range(size)]
    my_series = sblocks[rank][1]
    send = local_t.contiguous().view(-1)
    recv = send.new_empty(my_series * n_total)
    dist.all_to_all_single(
        recv, send,
        output_split_sizes=[my_series * c for c in counts],
        input_split_sizes=[sc * mine for _, sc in sblocks])
    out = send.new_empty((my_series, n_total))
    pos = 0
    for o, c in zip(offs, counts):
        out[:, o:o + c] = recv[pos:pos + my_series * c].view(my_series, c)
        pos += my_series * c
    return out


def gather_series(mine, n_series):
    """All-gathers the per-series results (lo / hi limits) of the series
    blocks dealt by :func:`exchange_series` back into (n_series,) order."""
    rank, size = world()
    if size == 1:
        return mine
    sblocks = [shard_range(n_series, r, size) for r in range(size)]
    width = max(c for _, c in sblocks)
    padded = mine.new_zeros((width,) + tuple(mine.shape[1:]))
    padded[:mine.shape[0]] = mine
    out = mine.new_empty((size * width,) + tuple(mine.shape[1:]))
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * width:r * width + c]
                      for r, (_, c) in enumerate(sblocks)], dim=0)


def percentile_sharded(eng, local, n_total, q_lo, q_hi):
    """
    Percentile limits over the resample axis of a per-resample output whose
    blocks live on different ranks (compute.boot_ci, pyls/compute.py:184-209).

    ``local`` (count_r, ...) is this rank's block.  Every rank selects the
    order statistics of ITS share of the series over all ``n_total``
    resamples; the (tiny) limits are all-gathered.  Returns (lo, hi) shaped
    like one resample's entry, identical on every rank.
    """
    _, size = world()
    if size == 1:
        return eng.percentile(local, q_lo, q_hi)
    tail = tuple(local.shape[1:])
    n_series = 1
    for t in tail:
        n_series *= int(t)
    local_t = eng.transpose(local.reshape(local.shape[0], n_series))
    mine = exchange_series(local_t, n_total)
    lo, hi = eng.percentile_series(mine, q_lo, q_hi)
    both = gather_series(torch.stack([lo, hi], dim=1), n_series)
    return (both[:, 0].reshape(tail).contiguous(),
            both[:, 1].reshape(tail).contiguous())


def broadcast_seed(rs):
    """An integer seed drawn from rank 0's RandomState `rs`, the same on every
    rank (analyses started with ``seed=None`` or a RandomState object)."""
    rank, size = world()
    box = [int(rs.randint(0, 2 ** 31 - 1)) if rank == 0 else None]
    if size > 1:
        dist.broadcast_object_list(box, src=0)
    return box[0]


def reduce_sum(*tensors):
    """In-place all-reduce (sum) of every tensor given."""
    _, size = world()
    if size > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors
