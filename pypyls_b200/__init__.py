# -*- coding: utf-8 -*-
"""
pypyls_b200 -- B200 (sm_100a) resampling engine for partial least squares
behind the front-end of netneurolab/pypyls: ``behavioral_pls`` and
``meancentered_pls`` returning ``PLSResults``; the permutation and bootstrap
loops run as batched CUDA launches through ``libplsb200.so``.
"""

__all__ = ['behavioral_pls', 'meancentered_pls', 'pls_regression', 'PLSResults', 'PLSInputs',
           'ResamplingEngine', 'release_workspaces', 'gen_permsamp', 'gen_bootsamp', 'save_results',
           'load_results', '__version__']

__version__ = '0.1.0'

from .structures import PLSInputs, PLSResults
from .resample import gen_bootsamp, gen_permsamp
from .engine import ResamplingEngine, release_workspaces
from .types import behavioral_pls, meancentered_pls, pls_regression
from .io import load_results, save_results
