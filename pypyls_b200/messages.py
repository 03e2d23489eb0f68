# -*- coding: utf-8 -*-
"""
User-visible error / warning texts of the reference front-end.

A drop-in replacement has to fail with the SAME messages (callers and the
reference's own tests match on them, e.g. pyls/tests/types/test_svd.py:128-143),
so the strings below are quoted verbatim from netneurolab/pypyls
(pyls/base.py:265-277, pyls/types/meancentered.py:18-36, GPL-2.0, (c) the pyls
developers); the checks that raise them are this package's own code.
"""

SAMPLES_MISMATCH = ('Number of samples specified by `groups` and '
                    '`n_cond` does not match number of samples in '
                    'input array(s).\n'
                    '    EXPECTED: {}\n'
                    '    ACTUAL:   {} (groups: {} * n_cond: {})')
XY_ROWS_DIFFER = ('Provided `X` and `Y` matrices must have the '
                  'same number of samples. Provided matrices '
                  'differed: X: {}, Y: {}')
NOT_DIVISIBLE = ('Provided `X` matrix with {} samples is not '
                 'evenly divisible into {} conditions. Please '
                 'confirm inputs are correct and try again. ')
ONE_GROUP_ONE_COND = ('Cannot perform PLS with only one group and one '
                      'condition. Please confirm inputs are correct.')
CENTERING_NEEDS_CONDITIONS = ('Cannot set mean_centering to 0 when there is only '
                              'one condition. Resetting mean_centering to 1.')
CENTERING_NEEDS_GROUPS = ('Cannot set mean_centering to 1 when there is only '
                          'one group. Resetting mean_centering to 0.')
BAD_CENTERING = 'Mean centering type must be in [0, 1, 2].'
