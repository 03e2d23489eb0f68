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
from ..engine import ResamplingEngine, to_host


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


class PLSRegression(BasePLS):
    def __init__(self, X, Y, *, n_components=None, n_perm=5000, n_boot=5000,
                 rotate=True, ci=95, aggfunc='mean', permsamples=None,
                 bootsamples=None, seed=None, verbose=True, n_proc=None,
                 **kwargs):
        X, Y = np.array(X, dtype=float), np.array(Y, dtype=float)
        if Y.ndim == 3:
            raise NotImplementedError(
                'Three-dimensional `Y` (aggfunc) is not part of the '
                'accelerated path yet.')
        if X.ndim != 2 or Y.ndim != 2:
            raise ValueError('`X` and `Y` must be two-dimensional arrays.')
        max_components = min(len(X) - 1, X.shape[1])
        if n_components is None:
            n_components = max_components
        else:
            n_components = int(n_components)
            if n_components > max_components:
                raise ValueError('Provided `n_components` cannot be greater '
                                 'than {}'.format(max_components))
        # rows that are missing altogether are masked like the reference does
        # (get_mask, pyls/types/regression.py:48-53); any other NaN fails there
        # inside sklearn's input validation with this message
        mask = get_mask(X, Y)
        if np.isnan(X[mask]).any() or np.isnan(Y[mask]).any():
            raise ValueError('Input contains NaN.')
        kwargs.update(n_split=0, test_split=0)
        super().__init__(X=X, Y=Y, n_components=n_components, n_perm=n_perm,
                         n_boot=n_boot, rotate=rotate, ci=ci, aggfunc=aggfunc,
                         permsamples=permsamples, bootsamples=bootsamples,
                         seed=seed, verbose=verbose, n_proc=n_proc, **kwargs)
        self.n_components = n_components
        self.results = self.run_pls(self.inputs.X, self.inputs.Y)

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
                               n_components=self.n_components)
        eng.set_data(X, Y)
        return eng

    def run_pls(self, X, Y):
        """Follows pyls/types/regression.py:375-428 and pyls/base.py:341-371."""
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
        res['x_weights'] = to_host(xw)
        res['x_scores'] = to_host(eng.project_scores(xw))
        res['x_scores'][np.isnan(X).any(axis=1)] = np.nan    # X @ x_weights
        varexp = pct.cpu().numpy()

        n_omega = max(self.inputs.n_perm, self.inputs.n_boot)
        self._omega = None
        if T > 11 and n_omega > 0:
            self._omega = eng.to_device(gaussian_tables(range(n_omega), T))

        if self.inputs.n_perm > 0:
            d_perm, _, _ = self.permutation(X, Y, seed=self.rs)
            res['permres']['pvals'] = eng.perm_pvals(
                self._dev['d_perm'], pct).cpu().numpy()
            res['permres']['permsamples'] = self.permsamp
            res['permres']['perm_singval'] = d_perm

        res['y_loadings'] = Y[mask].T @ res['x_scores'][mask]
        res['y_scores'] = np.full((len(Y), L), np.nan)
        res['y_scores'][mask] = resid_yscores(res['x_scores'][mask],
                                              Y[mask] @ res['y_loadings'])

        if self.inputs.n_boot > 0:
            distrib, u_sum, u_square = self.bootstrap(X, Y, self.rs)
            bsrs, uboot_se, corrci = self._boot_stats(add_orig=True)
            res['bootres'].update(dict(x_weights_normed=bsrs,
                                       x_weights_stderr=uboot_se,
                                       y_loadings=res['y_loadings'],
                                       y_loadings_boot=distrib,
                                       y_loadings_ci=corrci,
                                       bootsamples=self.bootsamp))
        res['varexp'] = varexp
        return res

    def _omega_block(self, first, count):
        return None if self._omega is None else \
            self._omega[first:first + count]

    def permutation(self, X, Y, seed=None):
        """Replaces BasePLS.permutation + PLSRegression._single_perm
        (pyls/types/regression.py:329-373): variance explained in Y per
        component for every permutation of the rows of Y."""
        n = self.inputs.n_perm
        table, block, first = self._table('perm', n, seed)
        local = self.engine.simpls_run_perms(
            block, self._omega_block(first, block.shape[0]))
        if isinstance(table, _DeviceTable):
            table.start()
        self.permsamp = _resolve(table)
        d_perm = pdist.gather_resamples(local, n)
        self._dev['d_perm'] = d_perm
        return to_host(d_perm).T.copy(), None, None

    def bootstrap(self, X, Y, seed=None):
        """Replaces BasePLS.bootstrap + PLSRegression._single_boot
        (pyls/types/regression.py:279-327)."""
        n = self.inputs.n_boot
        table, block, first = self._table('boot', n, seed)
        distrib, u_sum, u_square, _ = self.engine.simpls_run_boots(
            block, self._omega_block(first, block.shape[0]))
        if isinstance(table, _DeviceTable):
            table.start()
        self.bootsamp = _resolve(table)
        local = distrib
        distrib = pdist.gather_resamples(local, n)
        pdist.reduce_sum(u_sum, u_square)
        self._dev.update(distrib=distrib, distrib_local=local, u_sum=u_sum,
                         u_square=u_square)
        return (to_host(distrib.permute(1, 2, 0).contiguous()),
                to_host(u_sum), to_host(u_square))


def pls_regression(X, Y, *, n_components=None, n_perm=5000, n_boot=5000,
                   rotate=True, ci=95, aggfunc='mean', permsamples=None,
                   bootsamples=None, seed=None, verbose=True, n_proc=None,
                   **kwargs):
    """
    PLS regression (SIMPLS) of `Y` (S, T) on `X` (S, B); same call as
    ``pyls.pls_regression`` (pyls/types/regression.py:432-440) with the
    permutation test and bootstrap executed on the GPU.  Unlike the reference
    the caller's arrays are not centred in place.  Two-dimensional `Y` only;
    rows of `X` or `Y` that are missing altogether (all NaN) are masked as in
    the reference; ``n_proc`` is accepted but unused.

    Returns
    -------
    results : :obj:`pypyls_b200.structures.PLSResults`
    """
    pls = PLSRegression(X=X, Y=Y, n_components=n_components, n_perm=n_perm,
                        n_boot=n_boot, rotate=rotate, ci=ci, aggfunc=aggfunc,
                        permsamples=permsamples, bootsamples=bootsamples,
                        seed=seed, verbose=verbose, n_proc=n_proc, **kwargs)
    return pls.results
