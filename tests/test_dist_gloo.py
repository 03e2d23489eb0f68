# -*- coding: utf-8 -*-
"""
world_size-2 gloo test of the multi-GPU plumbing (pypyls_b200.dist) on CPU:
each rank computes its contiguous block of resamples with the CPU oracle, the
blocks are all-gathered / all-reduced, and the assembled result must equal the
single-process run.
"""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pls_oracle as po


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_perm, n_boot, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pypyls_b200 import dist as pdist
    rs = np.random.RandomState(0)
    groups, n_cond = [6, 7], 1
    X, Y = rs.rand(13, 30), rs.rand(13, 2)
    spec = po._Spec('behavioral', groups, n_cond)
    U, d, V = po.decompose(spec, X, Y, seed=1)
    ps = po.gen_permsamp(groups, n_cond, n_perm, seed=2)
    bs = po.gen_bootsamp(groups, n_cond, n_boot, seed=3)

    assert pdist.world() == (rank, world)
    first, count = pdist.my_block(n_perm)
    local = po.run_perms(spec, X, Y, ps, V, first=first, count=count)
    full = pdist.gather_resamples(torch.from_numpy(local.T.copy()), n_perm)

    first, count = pdist.my_block(n_boot)
    dl, us, uq = po.run_boots(spec, X, Y, bs, U, first=first, count=count)
    dfull = pdist.gather_resamples(
        torch.from_numpy(np.moveaxis(dl, -1, 0).copy()), n_boot)
    us, uq = torch.from_numpy(us), torch.from_numpy(uq)
    pdist.reduce_sum(us, uq)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), d_perm=full.numpy(),
             distrib=dfull.numpy(), u_sum=us.numpy(), u_square=uq.numpy())
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    n_perm, n_boot, world = 9, 7, 2       # odd counts -> ragged blocks
    mp.spawn(_worker, args=(world, _free_port(), n_perm, n_boot,
                            str(tmp_path)), nprocs=world, join=True)
    rs = np.random.RandomState(0)
    groups, n_cond = [6, 7], 1
    X, Y = rs.rand(13, 30), rs.rand(13, 2)
    spec = po._Spec('behavioral', groups, n_cond)
    U, d, V = po.decompose(spec, X, Y, seed=1)
    ps = po.gen_permsamp(groups, n_cond, n_perm, seed=2)
    bs = po.gen_bootsamp(groups, n_cond, n_boot, seed=3)
    want_p = po.run_perms(spec, X, Y, ps, V)
    want_d, want_us, want_uq = po.run_boots(spec, X, Y, bs, U)
    for rank in range(world):
        z = np.load(tmp_path / ('rank%d.npz' % rank))
        np.testing.assert_allclose(z['d_perm'], want_p.T, rtol=1e-12)
        np.testing.assert_allclose(z['distrib'], np.moveaxis(want_d, -1, 0),
                                   rtol=1e-12)
        np.testing.assert_allclose(z['u_sum'], want_us, rtol=1e-10)
        np.testing.assert_allclose(z['u_square'], want_uq, rtol=1e-10)


def _exchange_worker(rank, world, port, n_total, n_series, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pypyls_b200 import dist as pdist
    from pypyls_b200.resample import shard_range
    full = torch.arange(float(n_total * n_series)).reshape(n_total, n_series)
    first, count = pdist.my_block(n_total)
    local = full[first:first + count]
    # series-major blocks dealt over the ranks: rank q ends up with its share
    # of the series over ALL resamples, in resample-id order
    mine = pdist.exchange_series(local.t().contiguous(), n_total)
    s0, sc = shard_range(n_series, rank, world)
    assert torch.equal(mine, full[:, s0:s0 + sc].t())
    # per-series results back in series order on every rank
    got = pdist.gather_series(mine.sum(dim=1), n_series)
    assert torch.equal(got, full.sum(dim=0))
    # whole per-resample output on the root only
    root = pdist.gather_to_root(local, n_total)
    assert (root is None) == (rank != 0)
    if rank == 0:
        assert torch.equal(root, full)
    assert torch.equal(pdist.gather_resamples(local, n_total), full)
    open(os.path.join(out_dir, 'ok%d' % rank), 'w').close()
    dist.destroy_process_group()


def test_series_exchange_even_and_ragged(tmp_path):
    for n_total, n_series in ((8, 6), (9, 7), (5, 3)):
        out = tmp_path / ('%d_%d' % (n_total, n_series))
        out.mkdir()
        mp.spawn(_exchange_worker, args=(2, _free_port(), n_total, n_series,
                                         str(out)), nprocs=2, join=True)
        assert (out / 'ok0').exists() and (out / 'ok1').exists()


def test_single_process_is_a_no_op():
    from pypyls_b200 import dist as pdist
    t = torch.arange(6.).reshape(3, 2)
    assert pdist.world() == (0, 1)
    assert pdist.my_block(10) == (0, 10)
    assert pdist.gather_resamples(t, 3) is t
    assert pdist.reduce_sum(t)[0] is t


class _HostEngine:
    """Stands in for ResamplingEngine in the CPU test of broadcast_from_root."""
    device = torch.device('cpu')

    @staticmethod
    def to_device(a):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)


def _bcast_worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pypyls_b200 import dist as pdist
    # every rank holds different numbers: only rank 0's may come out
    X = np.full((5, 7), float(rank + 1)) + np.arange(7)
    got = pdist.broadcast_from_root(_HostEngine(), X, (5, 7))
    np.save(os.path.join(out_dir, 'bcast%d.npy' % rank), got.numpy())
    dist.destroy_process_group()


def test_inputs_broadcast_from_root(tmp_path):
    """input_source='root': rank 0 uploads, the other ranks receive its arrays."""
    world = 2
    mp.spawn(_bcast_worker, args=(world, _free_port(), str(tmp_path)),
             nprocs=world, join=True)
    want = np.full((5, 7), 1.0) + np.arange(7)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ('bcast%d.npy' % r)), want)
