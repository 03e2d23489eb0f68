# -*- coding: utf-8 -*-
"""
PLS regression (SIMPLS) front-end (interface of pyls/types/regression.py:
189-440) on the CUDA resampling engine.
"""

import numpy as np
import torch

from .. import dist as pdist
from .. import structures
from ..base import BasePLS, _DeviceTable, _resolve
from ..engine import Download, ResamplingEngine, copy_stream, to_host
from ..resample import gen_bootsamp


def resid_yscores(x_scores, y_scores):
    """Orthogonalises every column of `y_scores` against the preceding
    columns of `x_scores` (two Gram-Schmidt passes), as
    pyls/types/regression.py:9-45.  (S, L) host arrays."""
    x_scores = np.asarray(x_scores, dtype=float)
    out = np.array(y_scores, dtype=float)
    for comp in range(1, x_scores.shape[1]):
        col = out[:, comp]
        prev = x_scores[:, :comp]
        for _ in range(2):
            for j in range(comp):
                col = col - (prev[:, j] @ col) * prev[:, j]
        out[:, comp] = col
    return out


def get_mask(X, Y):
    """Rows where neither `X` nor `Y` is missing altogether (all NaN);
    pyls/types/regression.py:48-53."""
    return np.logical_not(np.logical_or(np.all(np.isnan(X), axis=1),
                                        np.all(np.isnan(Y), axis=1)))


def gaussian_tables(seeds, T):
    """(len(seeds), T, 11) test matrices: what sklearn's randomized_svd draws
    for ``compute.svd(Cov, n_components=1, seed=i)`` with an integer seed
    (pyls/types/regression.py:103 -> pyls/compute.py:48-49)."""
    out = np.empty((len(seeds), T, 11))
    rs = np.random.RandomState(0)
    for n, i in enumerate(seeds):
        rs.seed(int(i))     # same stream as RandomState(i), 15x cheaper
        out[n] = rs.standard_normal((T, 11))
    return out


AGGFUNCS = dict(mean=np.mean, median=np.median, sum=np.sum)


def _pair_table(s, c):
    """(2, n_boot) object array of (row sample, third-axis sample) pairs, the
    layout the reference keeps for a three-dimensional Y
    (pyls/types/regression.py:209-215)."""
    out = np.empty((2, s.shape[1]), dtype=object)
    for i in range(s.shape[1]):
        out[0, i], out[1, i] = s[:, i], c[:, i]
    return out


class PLSRegression(BasePLS):
    def __init__(self, X, Y, *, n_components=None, n_perm=5000, n_boot=5000,
                 rotate=True, ci=95, aggfunc='mean', permsamples=None,
                 bootsamples=None, seed=None, verbose=True, n_proc=None,
                 **kwargs):
        X, Y = np.array(X, dtype=float), np.array(Y, dtype=float)
        if X.ndim != 2 or Y.ndim not in (2, 3):
            raise ValueError('`X` must be two-dimensional and `Y` two- or '
                             'three-dimensional.')
        max_components = min(len(X) - 1, X.shape[1])
        if n_components is None:
            n_components = max_components
        else:
            n_components = int(n_components)
            if n_components > max_components:
                raise ValueError('Provided `n_components` cannot be greater '
                                 'than {}'.format(max_components))
        self._third = None
        if Y.ndim == 3:
            # every bootstrap also resamples the third axis of Y and aggregates
            # it (pyls/types/regression.py:207-235): the tables come in pairs
            S, C = Y.shape[0], Y.shape[-1]
            if bootsamples is None:
                rows = gen_bootsamp([S], 1, n_boot, seed=seed, verbose=False)
                third = gen_bootsamp([C], 1, n_boot, seed=seed, verbose=False)
            else:
                bs = np.asarray(bootsamples, dtype=object)
                ok = bs.shape == (2, n_boot)
                if ok:
                    rows = np.stack([np.asarray(b, dtype=int)
                                     for b in bs[0]], axis=-1)
                    third = np.stack([np.asarray(b, dtype=int)
                                      for b in bs[1]], axis=-1)
                    ok = rows.shape == (S, n_boot) and \
                        third.shape == (C, n_boot)
                if not ok:
                    raise ValueError('Provided bootsamples arrays does not '
                                     'match size of provided input arrays or '
                                     'number of bootstraps requested via '
                                     '`nboot`.')
            if not callable(aggfunc) and aggfunc not in AGGFUNCS:
                raise ValueError('Provided `aggfunc` must either be callable '
                                 'or one of {}'.format(sorted(AGGFUNCS)))
            self.aggfunc = AGGFUNCS.get(aggfunc, aggfunc)
            self._third = third
            bootsamples = rows
        # rows that are missing altogether are masked like the reference does
        # (get_mask, pyls/types/regression.py:48-53); any other NaN fails there
        # inside sklearn's input validation with this message
        Y2 = Y if Y.ndim == 2 else self._aggregate(Y)
        mask = get_mask(X, Y2)
        if np.isnan(X[mask]).any() or np.isnan(Y2[mask]).any():
            raise ValueError('Input contains NaN.')
        if Y.ndim == 3 and not mask.all():
            raise ValueError('Missing rows are not supported together with a '
                             'three-dimensional `Y`.')
        kwargs.update(n_split=0, test_split=0)
        super().__init__(X=X, Y=Y, n_components=n_components, n_perm=n_perm,
                         n_boot=n_boot, rotate=rotate, ci=ci, aggfunc=aggfunc,
                         permsamples=permsamples, bootsamples=bootsamples,
                         seed=seed, verbose=verbose, n_proc=n_proc, **kwargs)
        self.n_components = n_components
        self.results = self.run_pls(self.inputs.X, self.inputs.Y)

    def _aggregate(self, Y3):
        try:
            return self.aggfunc(Y3, axis=-1)
        except TypeError:
            raise TypeError('Provided callable `aggfun` must accept `axis` '
                            'keyword argument to condense an array along '
                            'the specified axis.')

    def engine_mode(self):
        return 'regression'

    def _make_engine(self, X, Y):
        device = self.inputs.get('device')
        if device is None:
            device = torch.cuda.current_device() \
                if torch.cuda.is_available() else 0
        eng = ResamplingEngine('regression', X.shape[0], X.shape[1],
                               Y.shape[1], [X.shape[0]], 1, device=device,
                               workspace_bytes=self.inputs.get(
                                   'workspace_bytes'),
                               n_components=self.n_components,
                               gemm_backend=self.inputs.get(
                                   'gemm_backend') or 'auto',
                               gemm_slices=self.inputs.get(
                                   'gemm_slices') or 6)
        if self.inputs.get('input_source') == 'root':
            from .. import dist as pdist
            X = pdist.broadcast_from_root(eng, X, X.shape)
            Y = pdist.broadcast_from_root(eng, Y, Y.shape)
        eng.set_data(X, Y)
        return eng

    def run_pls(self, X, Y):
        """Follows pyls/types/regression.py:375-428 and pyls/base.py:341-371.
        All device work is queued first (decomposition, permutations,
        bootstraps, statistics), every result rides to the host on the side
        stream behind its producer, and the host-side post-processing runs at
        the end while the later kernels are still busy."""
        Y3 = Y if Y.ndim == 3 else None
        if Y3 is not None:
            Y = self._aggregate(Y3)
        # the reference centres the caller's arrays in place; copies here
        X -= np.nanmean(X, axis=0, keepdims=True)
        Y -= np.nanmean(Y, axis=0, keepdims=True)
        mask = get_mask(X, Y)
        self.res = res = structures.PLSResults(inputs=self.inputs)
        self._later = []
        if mask.all():
            self.engine = eng = self._make_engine(X, Y)
        else:
            # missing rows travel zero-filled together with the row mask: every
            # decomposition then uses the rows of the resampled matrices whose X
            # and Y sources are both present (regression.py:271-272, 322-323)
            self.engine = eng = self._make_engine(np.nan_to_num(X),
                                                  np.nan_to_num(Y))
            eng.simpls_set_row_mask(~np.all(np.isnan(X), axis=1),
                                    ~np.all(np.isnan(Y), axis=1))
        T, L = Y.shape[1], self.n_components

        # the reference draws one (T, 11) Gaussian matrix per component from
        # the analysis' RandomState (compute.svd(..., seed=self.rs))
        omega0 = np.stack([self.rs.normal(size=(T, 11)) for _ in range(L)])
        xw, pct = eng.simpls_decompose(omega0 if T > 11 else None)
        self._dev = dict(U=xw, d=torch.ones_like(pct), pct=pct)
        dl = [Download(t) for t in (xw, eng.project_scores(xw), pct)]

        # Gaussian test matrices of the resamples (permutation i and bootstrap i
        # share RandomState(i), pyls/base.py:646-648, 502-507): replayed on the
        # device for this rank's blocks of resample ids
        self._omega, self._omega_first = None, 0
        if T > 11:
            blocks = [pdist.my_block(n) for n in (self.inputs.n_perm,
                                                  self.inputs.n_boot) if n > 0]
            if blocks:
                lo = min(f for f, _ in blocks)
                hi = max(f + c for f, c in blocks)
                self._omega_first = lo
                self._omega = eng.gen_gaussian_tables(lo, hi - lo)

        perm = None
        if self.inputs.n_perm > 0:
            perm = self._permutation_device(X, Y, self.rs)
            perm['pvals'] = Download(eng.perm_pvals(perm['d_perm'], pct))
            perm['dl'] = Download(perm['d_perm'])

        boot = stats = None
        if self.inputs.n_boot > 0:
            boot = self._bootstrap_device(X, Y3 if Y3 is not None else Y,
                                          self.rs)
            stats = [Download(t) for t in
                     self._boot_stats(add_orig=True, device=True)]

        def fill():
            res['x_weights'] = dl[0].get()
            res['x_scores'] = dl[1].get()
            res['x_scores'][np.isnan(X).any(axis=1)] = np.nan   # X @ x_weights
            if perm is not None:
                self.permsamp = _resolve(perm['table'])
                res['permres']['pvals'] = perm['pvals'].get()
                res['permres']['permsamples'] = self.permsamp
                res['permres']['perm_singval'] = perm['dl'].get().T.copy()
            res['y_loadings'] = Y[mask].T @ res['x_scores'][mask]
            res['y_scores'] = np.full((len(Y), L), np.nan)
            res['y_scores'][mask] = resid_yscores(res['x_scores'][mask],
                                                  Y[mask] @ res['y_loadings'])
            if boot is not None:
                self.bootsamp = _resolve(boot['table'])
                if self._third is not None:
                    self.bootsamp = _pair_table(self.bootsamp, self._third)
                res['bootres'].update(dict(
                    x_weights_normed=stats[0].get(),
                    x_weights_stderr=stats[1].get(),
                    y_loadings=res['y_loadings'],
                    y_loadings_boot=self._host_distrib(boot),
                    y_loadings_ci=stats[2].get(),
                    bootsamples=self.bootsamp))
            res['varexp'] = dl[2].get()
        self._later.append(fill)
        self._finalize()
        return res

    def _omega_block(self, first, count):
        if self._omega is None:
            return None
        k = first - self._omega_first
        return self._omega[k:k + count]

    def _permutation_device(self, X, Y, seed=None):
        """Queues the permutation test (replaces BasePLS.permutation +
        PLSRegression._single_perm, pyls/types/regression.py:329-373: variance
        explained in Y per component for every permutation of the rows of
        Y); returns device results and the table without waiting."""
        n = self.inputs.n_perm
        table, block, first = self._table('perm', n, seed)
        local = self.engine.simpls_run_perms(
            block, self._omega_block(first, int(block.shape[0])))
        if isinstance(table, _DeviceTable):
            table.start()
        d_perm = pdist.gather_resamples(local, n)
        self._dev['d_perm'] = d_perm
        return dict(d_perm=d_perm, table=table)

    def permutation(self, X, Y, seed=None):
        out = self._permutation_device(X, Y, seed)
        self.permsamp = _resolve(out['table'])
        return to_host(out['d_perm']).T.copy(), None, None

    def _bootstrap_device(self, X, Y, seed=None):
        """Queues the bootstrap (replaces BasePLS.bootstrap +
        PLSRegression._single_boot, pyls/types/regression.py:279-327).  A
        three-dimensional `Y` brings one aggregated behaviour matrix per
        bootstrap (regression.py:308-310), built on the host for this rank's
        block and uploaded."""
        n = self.inputs.n_boot
        table, block, first = self._table('boot', n, seed)
        count = int(block.shape[0])
        rank, size = pdist.world()
        root_only = size > 1 and self.inputs.gather_results == 'root'
        yres = None
        if Y.ndim == 3:
            third = self._third[:, first:first + count]
            if self.aggfunc in (np.mean, np.sum):
                # linear aggregation: one product with the multiplicities
                mult = np.zeros((Y.shape[-1], count))
                np.add.at(mult, (third, np.arange(count)[None, :]), 1.0)
                if self.aggfunc is np.mean:
                    mult /= Y.shape[-1]
                yres = np.moveaxis(Y @ mult, -1, 0)
            else:
                yres = np.stack([self._aggregate(Y[..., third[:, i]])
                                 for i in range(count)])
            yres = np.ascontiguousarray(yres)
        local, u_sum, u_square, _ = self.engine.simpls_run_boots(
            block, self._omega_block(first, count), yres=yres)
        if isinstance(table, _DeviceTable):
            table.start()
        main = torch.cuda.current_stream(self.engine.device)
        side = copy_stream(self.engine.device)
        ready = torch.cuda.Event()
        ready.record(main)
        want_host = not root_only or rank == 0
        host = torch.empty((n, self.engine.T, self.engine.L),
                           dtype=torch.float64, pin_memory=True) \
            if want_host else None
        host_done = torch.cuda.Event()
        distrib = local
        with torch.cuda.stream(side):
            side.wait_event(ready)
            if size > 1:
                distrib = pdist.gather_to_root(local, n) if root_only \
                    else pdist.gather_resamples(local, n)
            if want_host:
                host.copy_(distrib, non_blocking=True)
            host_done.record(side)
        if distrib is not None:
            distrib.record_stream(side)
        pdist.reduce_sum(u_sum, u_square)
        self._dev.update(distrib=distrib, distrib_local=local, u_sum=u_sum,
                         u_square=u_square)
        return dict(distrib=distrib, host=host, host_done=host_done,
                    u_sum=u_sum, u_square=u_square, table=table, keep=local)

    def bootstrap(self, X, Y, seed=None):
        out = self._bootstrap_device(X, Y, seed)
        us, uq = to_host(out['u_sum']), to_host(out['u_square'])
        self.bootsamp = _resolve(out['table'])
        return self._host_distrib(out), us, uq


def pls_regression(X, Y, *, n_components=None, n_perm=5000, n_boot=5000,
                   rotate=True, ci=95, aggfunc='mean', permsamples=None,
                   bootsamples=None, seed=None, verbose=True, n_proc=None,
                   **kwargs):
    """
    PLS regression (SIMPLS) of `Y` (S, T) on `X` (S, B); same call as
    ``pyls.pls_regression`` (pyls/types/regression.py:432-440) with the
    permutation test and bootstrap executed on the GPU.  Unlike the reference
    the caller's arrays are not centred in place.  `Y` may be three-dimensional
    (S, T, C): the analysis then runs on ``aggfunc(Y, axis=-1)`` and every
    bootstrap also resamples the third axis (``bootsamples``: a (2, n_boot)
    object array of (rows, third-axis) samples, as in the reference).  Rows of
    `X` or `Y` that are missing altogether (all NaN) are masked as in the
    reference (two-dimensional `Y` only); ``n_proc`` is accepted but unused.

    Returns
    -------
    results : :obj:`pypyls_b200.structures.PLSResults`
    """
    pls = PLSRegression(X=X, Y=Y, n_components=n_components, n_perm=n_perm,
                        n_boot=n_boot, rotate=rotate, ci=ci, aggfunc=aggfunc,
                        permsamples=permsamples, bootsamples=bootsamples,
                        seed=seed, verbose=verbose, n_proc=n_proc, **kwargs)
    return pls.results
