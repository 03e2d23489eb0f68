# -*- coding: utf-8 -*-
"""PLS types with the reference's front-end functions (pyls/types/)."""

__all__ = ['behavioral_pls', 'meancentered_pls', 'BehavioralPLS',
           'MeanCenteredPLS']

from .behavioral import BehavioralPLS, behavioral_pls
from .meancentered import MeanCenteredPLS, meancentered_pls
