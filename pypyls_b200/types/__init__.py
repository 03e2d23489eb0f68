# -*- coding: utf-8 -*-
"""PLS types with the reference's front-end functions (pyls/types/)."""

__all__ = ['behavioral_pls', 'meancentered_pls', 'pls_regression',
           'BehavioralPLS', 'MeanCenteredPLS', 'PLSRegression']

from .behavioral import BehavioralPLS, behavioral_pls
from .meancentered import MeanCenteredPLS, meancentered_pls
from .regression import PLSRegression, pls_regression
