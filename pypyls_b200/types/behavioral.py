# -*- coding: utf-8 -*-
"""
Behavioral PLS front-end (interface of pyls/types/behavioral.py:10-242) on the
CUDA resampling engine.
"""

import numpy as np

from ..base import BasePLS, _resolve
from ..engine import Download, to_host
from ..resample import gen_splits


class BehavioralPLS(BasePLS):
    def __init__(self, X, Y, *, groups=None, n_cond=1, n_perm=5000,
                 n_boot=5000, n_split=0, test_size=0.25, test_split=0,
                 covariance=False, rotate=True, ci=95, permsamples=None,
                 bootsamples=None, seed=None, verbose=True, n_proc=None,
                 **kwargs):
        X, Y = np.asarray(X), np.asarray(Y)
        if X.ndim != 2 or Y.ndim != 2:
            raise ValueError('`X` and `Y` must be two-dimensional arrays.')
        super().__init__(X=X, Y=Y, groups=groups, n_cond=n_cond,
                         n_perm=n_perm, n_boot=n_boot, n_split=n_split,
                         test_size=test_size, test_split=test_split,
                         covariance=covariance, rotate=rotate, ci=ci,
                         permsamples=permsamples, bootsamples=bootsamples,
                         seed=seed, verbose=verbose, n_proc=n_proc, **kwargs)
        self.results = self.run_pls(self.inputs.X, self.inputs.Y)

    def engine_mode(self):
        return 'behavioral_cov' if self.inputs.get('covariance') \
            else 'behavioral'

    def crossval(self, X, Y, groups=None, seed=None):
        """
        Cross-validation on the device (replaces pyls/types/behavioral.py:82-170):
        ``test_split`` train / test splits generated like the reference does
        (gen_splits, pyls/base.py:162-229, replaying its NumPy stream), every
        split decomposed and scored in one batched launch sequence.

        Returns
        -------
        r_scores, r2_scores : (T, C) numpy.ndarray
        """
        r, r2 = self._crossval_device(seed)
        return to_host(r).T.copy(), to_host(r2).T.copy()

    def _crossval_device(self, seed):
        splits = gen_splits(self.inputs.groups, self.inputs.n_cond,
                            self.inputs.test_split, seed=seed,
                            test_size=self.inputs.test_size)
        return self.engine.crossval(splits)

    def run_pls(self, X, Y):
        """Follows pyls/types/behavioral.py:172-227.  Device work is queued
        first (decomposition, permutations, bootstraps, cross-validation);
        downloads and host-side post-processing follow in one go."""
        res = super().run_pls(X, Y)
        eng = self.engine

        # y_loadings = per-cell xcorr(x_scores, Y): the bootstrap distribution
        # kernel evaluated on the identity resample (X @ U has unit-norm U)
        ident = np.arange(eng.S)[:, None]
        y_load = Download(eng.run_boots(ident)[0][0])

        boot = stats = None
        if self.inputs.n_boot > 0:
            boot = self._bootstrap_device(X, Y, self.rs)
            stats = [Download(t) for t in
                     self._boot_stats(add_orig=True, device=True)]

        # cross-validated prediction of Y (pyls/types/behavioral.py:217-219)
        cv = None
        if self.inputs.get('test_split') is not None and \
                (self.inputs.get('test_size') or 0) > 0:
            cv = [Download(t) for t in self._crossval_device(self.rs)]

        def fill():
            # y_scores: every cell's rows of Y times that cell's block of V
            cells = np.repeat(res['inputs']['groups'], res['inputs']['n_cond'])
            T = Y.shape[1]
            res['y_scores'] = np.vstack([
                y @ res['y_weights'][j * T:(j + 1) * T]
                for j, y in enumerate(np.split(Y, np.cumsum(cells)[:-1]))])
            res['y_loadings'] = y_load.get()
            if boot is not None:
                self.bootsamp = _resolve(boot['table'])
                res['bootres'].update(dict(
                    x_weights_normed=stats[0].get(),
                    x_weights_stderr=stats[1].get(),
                    y_loadings=res['y_loadings'].copy(),
                    y_loadings_boot=self._host_distrib(boot),
                    y_loadings_ci=stats[2].get(),
                    bootsamples=self.bootsamp))
            if cv is not None:
                res['cvres'].update(dict(pearson_r=cv[0].get().T.copy(),
                                         r_squared=cv[1].get().T.copy()))
            sq = np.diag(res['singvals']) ** 2
            res['varexp'] = sq / np.sum(sq)
            res['singvals'] = np.diag(res['singvals'])
        self._later.append(fill)
        self._finalize()
        return res


def behavioral_pls(X, Y, *, groups=None, n_cond=1, n_perm=5000, n_boot=5000,
                   n_split=0, test_size=0.25, test_split=0,
                   covariance=False, rotate=True, ci=95, permsamples=None,
                   bootsamples=None, seed=None, verbose=True, n_proc=None,
                   **kwargs):
    """
    Behavioral PLS of `X` (S, B) and `Y` (S, T); same call as
    ``pyls.behavioral_pls`` (pyls/types/behavioral.py:231-242) with the
    permutation test and bootstrap executed on the GPU.

    Differences from the reference front-end: ``test_split`` defaults to 0
    (the reference's default is 100; pass it to get ``cvres``) and ``n_proc``
    is accepted but unused (resamples run as one batched launch).  ``n_split``
    runs the split-half resampling on the device.  Analyses with more than 80
    latent variables (K = cells x behaviours), more than 80 rows per pair of
    split halves or per stacked train / test split run through generic
    (slower) kernels.  With more latent rows than features (K > B) the
    decomposition keeps L = B latent variables like ``compute.svd`` does and
    every resample is solved on the feature side; ``n_split`` / ``test_split``
    are not available in that orientation.
    Extra keywords: ``index_backend``, ``device``, ``workspace_bytes``.

    Returns
    -------
    results : :obj:`pypyls_b200.structures.PLSResults`
    """
    pls = BehavioralPLS(X=X, Y=Y, groups=groups, n_cond=n_cond,
                        n_perm=n_perm, n_boot=n_boot, n_split=n_split,
                        test_size=test_size, test_split=test_split,
                        covariance=covariance, rotate=rotate, ci=ci,
                        permsamples=permsamples, bootsamples=bootsamples,
                        seed=seed, verbose=verbose, n_proc=n_proc, **kwargs)
    return pls.results
