# -*- coding: utf-8 -*-
"""
Result / input containers with the key sets of the reference
(pyls/structures.py:146-351, pyls/utils.py:16-125): dictionaries with
attribute access that silently drop keys outside their whitelist and compare
equal when all non-empty entries agree to six decimals.
"""

import os

import numpy as np


def _is_empty(value):
    if value is None:
        return True
    try:
        return len(value.keys()) == 0
    except (AttributeError, TypeError):
        return False


def _filled(d):
    if not isinstance(d, dict):
        raise TypeError('Provided input must be type dict, not {}'
                        .format(type(d)))
    return {k for k, v in d.items() if not _is_empty(v)}


class ResDict(dict):
    """dict with attribute access restricted to the keys in ``allowed``."""

    allowed = []

    def __init__(self, **kwargs):
        super().__init__()
        for key, val in kwargs.items():
            self[key] = val

    def __setitem__(self, key, val):
        if key in type(self).allowed:
            super().__setitem__(key, val)

    def update(self, *args, **kwargs):
        for key, val in dict(*args, **kwargs).items():
            self[key] = val

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, val):
        self[key] = val

    def __dir__(self):
        return list(self.keys())

    def __str__(self):
        shown = [k for k in type(self).allowed if k in _filled(self)]
        return '{}({})'.format(type(self).__name__, ', '.join(shown))

    __repr__ = __str__

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return False
        if _filled(self) != _filled(other):
            return False
        for key, mine in self.items():
            theirs = other.get(key)
            if mine is None and theirs is None:
                continue
            if isinstance(mine, dict) and isinstance(theirs, dict):
                if mine != theirs:
                    return False
                continue
            try:
                np.testing.assert_array_almost_equal(mine, theirs)
            except (TypeError, AssertionError):
                return False
        return True

    def __ne__(self, other):
        return not self == other

    __hash__ = None


class PLSInputs(ResDict):
    """Inputs of an analysis (pyls/structures.py:146-172).  ``index_backend``,
    ``device``, ``workspace_bytes``, ``perm_path``, ``gather_results``, ``gemm_backend``,
    ``gemm_slices`` and ``input_source`` are additions of this engine."""

    allowed = [
        'X', 'Y', 'groups', 'n_cond', 'n_perm', 'n_boot', 'n_split',
        'test_split', 'test_size', 'mean_centering', 'covariance', 'rotate',
        'ci', 'seed', 'verbose', 'n_proc', 'bootsamples', 'permsamples',
        'method', 'n_components', 'aggfunc', 'permindices',
        'index_backend', 'device', 'workspace_bytes', 'perm_path',
        'gather_results', 'gemm_backend', 'gemm_slices', 'input_source',
    ]

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if self.get('n_split') == 0:
            self['n_split'] = None
        if self.get('test_split') == 0:
            self['test_split'] = None
        n_proc = self.get('n_proc')
        if n_proc is not None:
            if n_proc == 'max' or n_proc == -1:
                self['n_proc'] = os.cpu_count()
            elif n_proc < 0:
                self['n_proc'] = os.cpu_count() + 1 + n_proc
        ts = self.get('test_size')
        if ts is not None and (ts < 0 or ts >= 1):
            raise ValueError('test_size must be in [0, 1). Provided value: {}'
                             .format(ts))


class PLSBootResults(ResDict):
    allowed = [
        'x_weights_normed', 'x_weights_stderr', 'bootsamples',
        'y_loadings', 'y_loadings_boot', 'y_loadings_ci',
        'contrast', 'contrast_boot', 'contrast_ci'
    ]


class PLSPermResults(ResDict):
    allowed = ['pvals', 'permsamples', 'perm_singval']


class PLSSplitHalfResults(ResDict):
    allowed = [
        'ucorr', 'vcorr', 'ucorr_pvals', 'vcorr_pvals',
        'ucorr_uplim', 'vcorr_uplim', 'ucorr_lolim', 'vcorr_lolim'
    ]


class PLSCrossValidationResults(ResDict):
    allowed = ['pearson_r', 'r_squared']


class PLSResults(ResDict):
    """Results of an analysis (pyls/structures.py:198-245)."""

    allowed = [
        'x_weights', 'y_weights', 'x_scores', 'y_scores',
        'y_loadings', 'singvals', 'varexp',
        'permres', 'bootres', 'splitres', 'cvres', 'inputs'
    ]

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.inputs = PLSInputs(**kwargs.get('inputs', kwargs))
        self.bootres = PLSBootResults(**kwargs.get('bootres', kwargs))
        self.permres = PLSPermResults(**kwargs.get('permres', kwargs))
        self.splitres = PLSSplitHalfResults(**kwargs.get('splitres', kwargs))
        self.cvres = PLSCrossValidationResults(**kwargs.get('cvres', kwargs))
