# -*- coding: utf-8 -*-
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


def load_golden(name):
    """Returns (inputs, outputs) dicts of a tests/golden/<name>.npz fixture."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    ins = {k[3:]: z[k] for k in z.files if k.startswith('in_')}
    outs = {k[4:]: z[k] for k in z.files if k.startswith('out_')}
    for k in ('n_cond', 'n_perm', 'n_boot', 'seed', 'mean_centering',
              'n_components', 'ci', 'n_split', 'test_split'):
        if k in ins:
            ins[k] = int(ins[k])
    for k in ('rotate', 'covariance'):
        if k in ins:
            ins[k] = bool(ins[k])
    if 'groups' in ins:
        ins['groups'] = [int(g) for g in np.atleast_1d(ins['groups'])]
    return ins, outs


@pytest.fixture(scope='session')
def golden():
    return load_golden
