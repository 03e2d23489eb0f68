# -*- coding: utf-8 -*-
"""
Parity at benchmark scale: every BASELINE configuration through the public
front-end, compared

* element-wise with the CPU oracle on 200 permutations + 200 bootstraps (the
  oracle runs in a pool of worker processes with single-threaded BLAS, the
  reference's own parallel mode, pyls/utils.py:252-279), and
* at FULL n (device-generated tables) with the vectorised fast oracle for the
  quantities north_star names -- permutation p-values and bootstrap CIs.

Tolerances: per-resample values 1e-8 relative (1e-7 for bootstrap ratios),
p-values exact, CIs 1e-9 absolute; north_star asks for 1e-5.
"""

import multiprocessing as mp
import os

import numpy as np
import pytest

from oracle import pls_oracle as po
from test_gpu_parity import check_rank_deficient_bsr, close

pytestmark = pytest.mark.gpu

N_EACH = 200
_W = {}


def _make(case):
    rs = np.random.RandomState(1234)
    kind = case['kind']
    X, Y = rs.rand(case['S'], case['B']), rs.rand(case['S'], case['T'])
    if case.get('planted'):
        # low-rank signal variant of config 4
        Y[:, :3] += X[:, :50] @ rs.rand(50, 3) * 0.15
    if kind == 'regression':
        X, Y = X - X.mean(0), Y - Y.mean(0)
        spec = po._Spec('regression', [case['S']], 1, n_components=case['L'])
    elif kind == 'meancentered':
        spec = po._Spec('meancentered', case['groups'], case['n_cond'])
        Y = spec.dummy
    else:
        spec = po._Spec('behavioral', case['groups'], case['n_cond'])
    return spec, X, Y


def _init():
    try:
        from threadpoolctl import threadpool_limits
        _W['limit'] = threadpool_limits(1)
    except Exception:
        pass
    import warnings
    warnings.filterwarnings('ignore')


def _task(args):
    case, uv_path, kind, first, cols = args
    key = repr(sorted(case.items()))
    if _W.get('key') != key:            # one data set resident per worker
        spec, X, Y = _make(case)
        _W.update(key=key, spec=spec, X=X, Y=Y, uv=None)
    if _W.get('uv') != uv_path:
        z = np.load(uv_path)
        _W.update(uv=uv_path, U=z['U'], V=z['V'] if 'V' in z.files else None)
    w = _W
    if kind == 'perm':
        # resample i is seeded with its number (pyls/base.py:646-648)
        return first, np.stack([
            po.single_perm(w['spec'], w['X'], w['Y'], cols[:, j], w['V'],
                           seed=first + j) for j in range(cols.shape[1])], -1)
    dist, us, uq = [], 0.0, 0.0
    for j in range(cols.shape[1]):
        d, u = po.single_boot(w['spec'], w['X'], w['Y'], cols[:, j], w['U'],
                              seed=first + j)
        dist.append(d)
        us, uq = us + u, uq + u ** 2
    return first, np.stack(dist, -1), us, uq


@pytest.fixture(scope='module')
def oracle_pool():
    """One pool of oracle workers for the whole module (spawning 16 Python
    processes that import scipy / scikit-learn costs more than the work)."""
    workers = max(1, min(len(os.sched_getaffinity(0)), 32))
    with mp.get_context('spawn').Pool(workers, initializer=_init) as pool:
        pool.workers = workers
        yield pool


def pool_oracle(pool, tmp_path, case, U, V, ps, bs):
    """(d_perm (L, n), distrib (., L, n), u_sum, u_square) of the oracle over
    the tables, computed by the pool."""
    uv = str(tmp_path / 'uv.npz')
    np.savez(uv, **({'U': U} if V is None else {'U': U, 'V': V}))
    per = max(1, ps.shape[1] // (2 * pool.workers))
    tasks = [(case, uv, 'perm', a, ps[:, a:a + per])
             for a in range(0, ps.shape[1], per)]
    tasks += [(case, uv, 'boot', a, bs[:, a:a + per])
              for a in range(0, bs.shape[1], per)]
    done = pool.map(_task, tasks, chunksize=1)
    perms = sorted((r for r in done if len(r) == 2), key=lambda r: r[0])
    boots = sorted((r for r in done if len(r) == 4), key=lambda r: r[0])
    d_perm = np.concatenate([r[1] for r in perms], axis=-1)
    distrib = np.concatenate([r[1] for r in boots], axis=-1)
    us = sum(r[2] for r in boots)
    uq = sum(r[3] for r in boots)
    return d_perm, distrib, us, uq


CFG2 = dict(kind='behavioral', S=80, B=10000, T=10, groups=[20, 20], n_cond=2,
            n_perm=5000, n_boot=5000)
CFG3 = dict(kind='meancentered', S=320, B=20000, T=1, groups=[40] * 4,
            n_cond=2, n_perm=10000, n_boot=10000)
CFG5 = dict(kind='behavioral', S=200, B=100000, T=10, groups=[200], n_cond=1,
            n_perm=10000, n_boot=10000)
CFG4 = dict(kind='regression', S=500, B=5000, T=20, L=10, groups=[500],
            n_cond=1)


def _front_end(case, X, Y, **kw):
    import pypyls_b200 as pyls
    if case['kind'] == 'meancentered':
        return pyls.meancentered_pls(X, groups=case['groups'],
                                     n_cond=case['n_cond'], verbose=False,
                                     **kw)
    return pyls.behavioral_pls(X, Y, groups=case['groups'],
                               n_cond=case['n_cond'], verbose=False, **kw)


@pytest.mark.parametrize('case', [CFG2, CFG3, CFG5],
                         ids=['cfg2', 'cfg3', 'cfg5'])
def test_baseline_config_matches_oracle_on_200_plus_200(case, oracle_pool,
                                                        tmp_path):
    spec, X, Y = _make(case)
    ps = po.gen_permsamp(case['groups'], case['n_cond'], N_EACH, seed=1)
    bs = po.gen_bootsamp(case['groups'], case['n_cond'], N_EACH, seed=2)
    out = _front_end(case, X, Y, n_perm=N_EACH, n_boot=N_EACH, seed=1234,
                     permsamples=ps, bootsamples=bs)
    U, V = out.x_weights, out.y_weights
    sv = out.singvals
    live = sv > 1e-8 * sv.max()
    d_perm, distrib, us, uq = pool_oracle(oracle_pool, tmp_path, case, U, V,
                                          ps, bs)
    close(out.permres.perm_singval[live], d_perm[live])
    pv = po.perm_sig(np.diag(sv), d_perm)
    assert np.array_equal(out.permres.pvals[live], pv[live])
    mc = case['kind'] == 'meancentered'
    boot = out.bootres.contrast_boot if mc else out.bootres.y_loadings_boot
    ci = out.bootres.contrast_ci if mc else out.bootres.y_loadings_ci
    close(boot[:, live], distrib[:, live], atol=1e-10)
    close(ci[:, live], np.stack(po.boot_ci(distrib), -1)[:, live], atol=1e-10)
    # bootstrap ratios: the reference's formula on the oracle's accumulators
    bs_orig = U * sv[None, :]
    if mc:
        want, _ = po.boot_rel(bs_orig, us, uq, N_EACH)
    else:
        want, _ = po.boot_rel(bs_orig, us + bs_orig, uq + bs_orig ** 2,
                              N_EACH + 1)
    if case is CFG5:
        # K = 10 of 200 subjects: every bootstrap is full rank -> the reference's
        # own rule, tightly
        close(out.bootres.x_weights_normed, want, rtol=1e-7, atol=1e-9)
    else:
        # rank-deficient resamples (mean-centred; 20-subject cells with T = 10):
        # null-safe rule tightly, the reference within its own noise
        check_rank_deficient_bsr(out, X, None if mc else Y, want, live)


@pytest.mark.parametrize('case', [CFG2, CFG3, CFG5],
                         ids=['cfg2', 'cfg3', 'cfg5'])
def test_baseline_config_full_n_pvalues_and_cis(case):
    """The whole job with device-generated tables; p-values and CIs against
    the vectorised fast oracle on the very same tables."""
    spec, X, Y = _make(case)
    out = _front_end(case, X, Y, n_perm=case['n_perm'], n_boot=case['n_boot'],
                     seed=1234)
    sv = out.singvals
    live = sv > 1e-8 * sv.max()
    fast = po.fast_stats(spec, X, Y, out.permres.permsamples,
                         out.bootres.bootsamples, out.x_weights, sv,
                         out.y_weights)
    close(out.permres.perm_singval[live], fast['perm_singval'][live])
    assert np.array_equal(out.permres.pvals[live], fast['pvals'][live])
    mc = case['kind'] == 'meancentered'
    boot = out.bootres.contrast_boot if mc else out.bootres.y_loadings_boot
    ci = out.bootres.contrast_ci if mc else out.bootres.y_loadings_ci
    close(boot[:, live], fast['distrib'][:, live], atol=1e-10)
    np.testing.assert_allclose(ci[:, live], fast['distrib_ci'][:, live],
                               rtol=0, atol=1e-9)


@pytest.mark.parametrize('planted', [False, True], ids=['noise', 'planted'])
def test_config4_simpls_matches_oracle_on_200_plus_200(planted, oracle_pool,
                                                       tmp_path):
    """pls_regression at config-4 shape, pure noise (where the reference's
    randomized top-1 SVD is most seed-sensitive) and with a planted
    low-rank signal: the engine reproduces the SAME randomized subspace."""
    import pypyls_b200 as pyls
    case = dict(CFG4, planted=planted)
    spec, X, Y = _make(case)
    n_each = 64 if planted else N_EACH
    ps = po.gen_permsamp([case['S']], 1, n_each, seed=1)
    bs = po.gen_bootsamp([case['S']], 1, n_each, seed=2)
    out = pyls.pls_regression(X, Y, n_components=case['L'], n_perm=n_each,
                              n_boot=n_each, seed=1234, verbose=False,
                              permsamples=ps, bootsamples=bs)
    W = out.x_weights
    d_perm, distrib, us, uq = pool_oracle(oracle_pool, tmp_path, case, W,
                                          None, ps, bs)
    close(out.permres.perm_singval, d_perm, rtol=1e-7)
    pv = po.perm_sig(np.diag(out.varexp), d_perm)
    assert np.array_equal(out.permres.pvals, pv)
    close(out.bootres.y_loadings_boot, distrib, rtol=1e-6, atol=1e-9)
    want, _ = po.boot_rel(W, us + W, uq + W ** 2, n_each + 1)
    close(out.bootres.x_weights_normed, want, rtol=1e-6, atol=1e-8)
