# -*- coding: utf-8 -*-
"""
Mean-centered PLS front-end (interface of pyls/types/meancentered.py:10-195)
on the CUDA resampling engine.
"""

import warnings

import numpy as np

from .. import messages
from ..base import BasePLS, _resolve, normalise_groups
from ..engine import Download


def dummy_code(groups, n_cond=1):
    """(S, J) 0/1 membership of every row in its group x condition cell
    (pyls/utils.py:155-175)."""
    sizes = np.repeat([int(g) for g in groups], n_cond)
    labels = np.repeat(np.arange(len(sizes)), sizes)
    return (labels[:, None] == np.arange(len(sizes))[None, :]).astype(int)


class MeanCenteredPLS(BasePLS):
    def __init__(self, X, groups=None, n_cond=1, mean_centering=0, n_perm=5000,
                 n_boot=5000, n_split=0, rotate=True, ci=95,
                 permsamples=None, bootsamples=None, seed=None,
                 verbose=True, n_proc=None, **kwargs):
        X = np.asarray(X)
        if groups is None and len(X) % n_cond:
            raise ValueError(messages.NOT_DIVISIBLE.format(len(X), n_cond))
        groups = normalise_groups(groups, len(X), n_cond)
        one_group, one_cond = len(groups) == 1, n_cond == 1
        if one_group and one_cond:
            raise ValueError(messages.ONE_GROUP_ONE_COND)
        # a centring that has nothing to average over falls back to the other
        # one, with the reference's warning (pyls/types/meancentered.py:29-36)
        if one_cond and mean_centering == 0:
            warnings.warn(messages.CENTERING_NEEDS_CONDITIONS)
            mean_centering = 1
        elif one_group and mean_centering == 1:
            warnings.warn(messages.CENTERING_NEEDS_GROUPS)
            mean_centering = 0
        if mean_centering not in (0, 1, 2):
            raise ValueError(messages.BAD_CENTERING)

        super().__init__(X=X, groups=groups, n_cond=n_cond,
                         mean_centering=mean_centering, n_perm=n_perm,
                         n_boot=n_boot, n_split=n_split, rotate=rotate, ci=ci,
                         permsamples=permsamples, bootsamples=bootsamples,
                         seed=seed, verbose=verbose, n_proc=n_proc, **kwargs)
        self.inputs.Y = dummy_code(self.inputs.groups, self.inputs.n_cond)
        self.results = self.run_pls(self.inputs.X, self.inputs.Y)

    def engine_mode(self):
        return 'meancentered'

    def _engine_y(self, Y):
        return None

    def run_pls(self, X, Y):
        """Follows pyls/types/meancentered.py:127-179.  Device work is queued
        first; downloads and host-side post-processing follow in one go."""
        res = super().run_pls(X, Y)
        eng = self.engine

        # contrast = cell means of the de-meaned rows projected on U: the
        # bootstrap distribution kernel evaluated on the identity resample
        ident = np.arange(eng.S)[:, None]
        contrast = Download(eng.run_boots(ident)[0][0])

        boot = stats = None
        if self.inputs.n_boot > 0:
            boot = self._bootstrap_device(X, Y, self.rs)
            stats = [Download(t) for t in
                     self._boot_stats(add_orig=False, device=True)]

        def fill():
            res['y_scores'] = Y @ res['y_weights']
            if boot is not None:
                self.bootsamp = _resolve(boot['table'])
                res['bootres'].update(dict(
                    x_weights_normed=stats[0].get(),
                    x_weights_stderr=stats[1].get(),
                    bootsamples=self.bootsamp,
                    contrast=contrast.get(),
                    contrast_boot=self._host_distrib(boot),
                    contrast_ci=stats[2].get()))
            sq = np.diag(res['singvals']) ** 2
            res['varexp'] = sq / np.sum(sq)
            res['singvals'] = np.diag(res['singvals'])
        self._later.append(fill)
        self._finalize()
        return res


def meancentered_pls(X, *, groups=None, n_cond=1, mean_centering=0,
                     n_perm=5000, n_boot=5000, n_split=0, rotate=True, ci=95,
                     permsamples=None, bootsamples=None, seed=None,
                     verbose=True, n_proc=None, **kwargs):
    """
    Mean-centered PLS of `X` (S, B) sorted into `groups` and conditions; same
    call as ``pyls.meancentered_pls`` (pyls/types/meancentered.py:182-195)
    with the permutation test and bootstrap executed on the GPU.
    ``n_split`` runs the split-half resampling on the device; ``n_proc`` is
    accepted but unused.

    Returns
    -------
    results : :obj:`pypyls_b200.structures.PLSResults`
    """
    pls = MeanCenteredPLS(X=X, groups=groups, n_cond=n_cond,
                          mean_centering=mean_centering,
                          n_perm=n_perm, n_boot=n_boot, n_split=n_split,
                          rotate=rotate, ci=ci, permsamples=permsamples,
                          bootsamples=bootsamples, seed=seed, verbose=verbose,
                          n_proc=n_proc, **kwargs)
    return pls.results
