# -*- coding: utf-8 -*-
"""
Two ranks on two GPUs (NCCL) through the real front-end: every rank runs its
block of the resamples, the results are all-gathered / all-reduced, and every
rank must hold the full result of the single-process analysis -- checked
against the CPU oracle, for a behavioural analysis with split-half resampling
and a mean-centred one.  Skipped on a box with fewer than two GPUs (the GPU
tier of the driver runs on one; `gpurun --gpus 2` runs it).
"""

import os
import socket

import numpy as np
import pytest
import torch

from oracle import pls_oracle as po

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _case():
    rs = np.random.RandomState(7)
    groups, n_cond = [9, 8], 2
    X, Y = rs.rand(34, 700), rs.rand(34, 3)
    Y[:, 0] += X[:, :30].mean(axis=1) * 2
    kw = dict(groups=groups, n_cond=n_cond, n_perm=11, n_boot=9, n_split=3,
              seed=21, permsamples=po.gen_permsamp(groups, n_cond, 11, seed=1),
              bootsamples=po.gen_bootsamp(groups, n_cond, 9, seed=2))
    return X, Y, kw


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world)
    import pypyls_b200 as pyls
    X, Y, kw = _case()
    out = pyls.behavioral_pls(X, Y, index_backend='reference', verbose=False,
                              device=rank, **kw)
    kwm = {k: v for k, v in kw.items() if k != 'n_split'}
    outm = pyls.meancentered_pls(X, verbose=False, device=rank,
                                 mean_centering=0, **kwm)
    # per-resample arrays on the host of rank 0 only; the statistics (p-values,
    # intervals from the series-sharded percentile exchange) everywhere
    outr = pyls.behavioral_pls(X, Y, index_backend='reference', verbose=False,
                               device=rank, gather_results='root',
                               **{k: v for k, v in kw.items()
                                  if k != 'n_split'})
    assert (outr.bootres.y_loadings_boot is None) == (rank != 0)
    if rank == 0:
        assert np.array_equal(outr.bootres.y_loadings_boot,
                              out.bootres.y_loadings_boot)
    assert np.array_equal(outr.bootres.y_loadings_ci,
                          out.bootres.y_loadings_ci)
    assert np.array_equal(outr.permres.pvals, out.permres.pvals)
    # input_source='root': the other ranks' own arrays only give the shapes
    Xr, Yr = (X, Y) if rank == 0 else (np.zeros_like(X), np.zeros_like(Y))
    outb = pyls.behavioral_pls(Xr, Yr, index_backend='reference', verbose=False,
                               device=rank, input_source='root',
                               **{k: v for k, v in kw.items()
                                  if k != 'n_split'})
    assert np.array_equal(outb.permres.perm_singval, out.permres.perm_singval)
    assert np.array_equal(outb.bootres.y_loadings_ci, out.bootres.y_loadings_ci)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank),
             perm=out.permres.perm_singval, pvals=out.permres.pvals,
             boot=out.bootres.y_loadings_boot, bsr=out.bootres.x_weights_normed,
             ci=out.bootres.y_loadings_ci, mci=outm.bootres.contrast_ci,
             ucorr_pvals=out.splitres.ucorr_pvals,
             ucorr_uplim=out.splitres.ucorr_uplim,
             mperm=outm.permres.perm_singval, mboot=outm.bootres.contrast_boot,
             msing=outm.singvals)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_front_end_matches_oracle(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world,
             join=True)
    X, Y, kw = _case()
    ref = po.behavioral_pls(X, Y, **kw)
    kwm = {k: v for k, v in kw.items() if k != 'n_split'}
    refm = po.meancentered_pls(X, mean_centering=0, **kwm)
    keep = ~np.isclose(refm['singvals'], 0)
    for rank in range(world):
        z = np.load(tmp_path / ('rank%d.npz' % rank))
        np.testing.assert_allclose(z['perm'], ref['perm_singval'], rtol=1e-8)
        assert np.array_equal(z['pvals'], ref['pvals'])
        np.testing.assert_allclose(z['boot'], ref['distrib'], rtol=1e-8,
                                   atol=1e-11)
        np.testing.assert_allclose(z['bsr'], ref['x_weights_normed'],
                                   rtol=1e-6)
        # percentile limits: every rank selected its share of the series
        np.testing.assert_allclose(z['ci'], ref['distrib_ci'], rtol=1e-8,
                                   atol=1e-11)
        np.testing.assert_allclose(z['mci'][:, keep],
                                   refm['distrib_ci'][:, keep], rtol=1e-8,
                                   atol=1e-11)
        assert np.array_equal(z['ucorr_pvals'], ref['ucorr_pvals'])
        np.testing.assert_allclose(z['ucorr_uplim'], ref['ucorr_uplim'],
                                   rtol=0, atol=1e-7)
        np.testing.assert_allclose(z['mperm'][keep], refm['perm_singval'][keep],
                                   rtol=1e-8)
        np.testing.assert_allclose(z['mboot'][:, keep],
                                   refm['distrib'][:, keep], rtol=1e-8,
                                   atol=1e-11)
