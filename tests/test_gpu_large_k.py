# -*- coding: utf-8 -*-
"""
The reference's own integration shapes (pyls/tests/types/test_svd.py:6-19,
67-100: X = rand(100, 1000), Y = rand(100, 100), one / several groups and
conditions, rotate True / False, n_perm = 20, n_boot = 10) through the CUDA
front-end, against the CPU oracle.  These have K = J * T = 100 ... 400 latent
variables -- beyond the fragment-table kernels (K <= 80), so they exercise the
generic tiled passes (csrc/large_k.cu) and the small-matrix kernels running out
of a global work space.

The reference asserts shapes only for these runs; here the values are compared
too.  A cell of n subjects gives a cross-correlation block of rank <= n - 1, so
most of the K latent variables are numerically null (singular value ~ 1e-15 in
the reference, arbitrary vectors): comparisons are restricted to the live ones,
as the reference's own Matlab harness does (pyls/tests/matlab.py:160).
"""

import numpy as np
import pytest

from oracle import pls_oracle as po
from test_gpu_parity import check_rank_deficient_bsr, close

pytestmark = pytest.mark.gpu

SUBJ, XF, YF = 100, 1000, 100
LAYOUTS = {
    'onegroup_onecond': ([100], 1),
    'multigroup_onecond': ([33, 34, 33], 1),
    'onegroup_multicond': ([25], 4),
    'multigroup_multicond': ([25, 25], 2),
}


# (layout, rotate, n_perm, n_boot): the reference runs 20 + 10 everywhere; the CPU
# oracle needs ~50 s for that at K = 400, so the largest layouts run fewer
CASES = [('onegroup_onecond', True, 20, 10), ('onegroup_onecond', False, 20, 10),
         ('multigroup_onecond', True, 6, 4), ('onegroup_multicond', False, 6, 4),
         ('multigroup_multicond', True, 6, 4)]


@pytest.mark.parametrize('layout,rotate,n_perm,n_boot', CASES)
def test_reference_integration_shapes(layout, rotate, n_perm, n_boot):
    import pypyls_b200 as pyls
    groups, n_cond = LAYOUTS[layout]
    rs = np.random.RandomState(1234)
    X, Y = rs.rand(SUBJ, XF), rs.rand(SUBJ, YF)
    K = len(groups) * n_cond * YF
    ps = po.gen_permsamp(groups, n_cond, n_perm, seed=1)
    bs = po.gen_bootsamp(groups, n_cond, n_boot, seed=2)
    kw = dict(groups=groups, n_cond=n_cond, n_perm=n_perm, n_boot=n_boot,
              seed=3, rotate=rotate, permsamples=ps, bootsamples=bs)
    out = pyls.behavioral_pls(X, Y, verbose=False, **kw)
    ref = po.behavioral_pls(X, Y, **kw)

    # shapes the reference's test asserts (test_svd.py:30-66)
    assert out.x_weights.shape == (XF, K) and out.y_weights.shape == (K, K)
    assert out.singvals.shape == (K,) and out.x_scores.shape == (SUBJ, K)
    assert out.permres.perm_singval.shape == (K, n_perm)
    assert out.bootres.x_weights_normed.shape == (XF, K)
    assert out.bootres.y_loadings_boot.shape == (K, K, n_boot)
    assert out.bootres.y_loadings_ci.shape == (K, K, 2)

    sv = ref['singvals']
    live = sv > 1e-8 * sv.max()
    assert 20 < live.sum() <= min(K, SUBJ)
    close(out.singvals[live], sv[live])
    assert np.all(out.singvals[~live] < 1e-6 * sv.max())
    close(out.x_weights[:, live], ref['x_weights'][:, live], rtol=1e-6,
          atol=1e-9)
    close(out.y_weights[:, live], ref['y_weights'][:, live], rtol=1e-6,
          atol=1e-9)
    if rotate:
        close(out.permres.perm_singval[live], ref['perm_singval'][live])
        assert np.array_equal(out.permres.pvals[live], ref['pvals'][live])
    else:
        # un-rotated: the singular values of every permuted cross-correlation
        top = ref['perm_singval'] > 1e-8 * ref['perm_singval'].max()
        close(out.permres.perm_singval[top], ref['perm_singval'][top],
              rtol=1e-7)
    # the bootstrap distribution needs the original weights only
    close(out.bootres.y_loadings_boot[:, live], ref['distrib'][:, live],
          rtol=1e-6, atol=1e-9)
    close(out.bootres.y_loadings_ci[:, live], ref['distrib_ci'][:, live],
          rtol=1e-6, atol=1e-9)
    # every resampled block is rank deficient: the ratios follow the null-safe rule
    check_rank_deficient_bsr(out, X, Y, ref['x_weights_normed'], live,
                             min_corr=-1.0, max_dev=np.inf)


def test_reference_integration_shape_with_split_half():
    """n_split = 5 at K = 100 (test_svd.py:67-75): 2 K = 200 rows per pair of
    halves, far beyond the fragment tables; shapes as the reference asserts
    them, values against the oracle's own split-half restatement."""
    import pypyls_b200 as pyls
    rs = np.random.RandomState(1234)
    X, Y = rs.rand(SUBJ, XF), rs.rand(SUBJ, YF)
    kw = dict(n_perm=4, n_boot=0, n_split=5, seed=3, rotate=True)
    out = pyls.behavioral_pls(X, Y, index_backend='reference', verbose=False,
                              **kw)
    ref = po.behavioral_pls(X, Y, **kw)
    assert out.splitres.ucorr.shape == (100,)
    assert out.splitres.ucorr_pvals.shape == (100,)
    sv = ref['singvals']
    live = sv > 1e-8 * sv.max()
    np.testing.assert_allclose(out.splitres.ucorr[live], ref['ucorr'][live],
                               rtol=0, atol=1e-7)
    np.testing.assert_allclose(out.splitres.vcorr[live], ref['vcorr'][live],
                               rtol=0, atol=1e-7)
