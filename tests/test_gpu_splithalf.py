# -*- coding: utf-8 -*-
"""
Split-half resampling on the device (n_split; BasePLS.split_half and its call
inside every permutation, pyls/base.py:373-397, 704-708, 714-770):
front-end -> ctypes -> plsb_split_half / plsb_gen_split_masks, against the
reference's own outputs (tests/golden, identical seeds), the CPU oracle, and
size-independent properties at BASELINE config-2 size.

Tolerances: the correlations are held to 1e-7 absolute (they come out of K x K
Gram matrices: eps * cond(R)^2), p-values exactly, percentile limits 1e-7.
Numerically null latent variables (mean-centred PLS always has some) are
rounding noise in the reference itself and are left out.
"""

import numpy as np
import pytest

from conftest import load_golden
from oracle import pls_oracle as po

pytestmark = pytest.mark.gpu

SPLIT_KEYS = ('ucorr', 'vcorr', 'ucorr_lolim', 'ucorr_uplim', 'vcorr_lolim',
              'vcorr_uplim')
ATOL = 1e-7


def _non_null(singvals):
    return singvals > 1e-8 * singvals.max()


@pytest.mark.parametrize('name', ['bpls_split_rot', 'bpls_split_cov_norot',
                                  'bpls_split_3g3c', 'bpls_split_prepermuted',
                                  'mpls_split_mc0', 'mpls_split_mc1',
                                  'mpls_split_mc2'])
def test_split_half_matches_reference_golden(name):
    import pypyls_b200 as pyls
    ins, ref = load_golden(name)
    X = ins.pop('X')
    if 'Yperm' in ins:
        out = pyls.behavioral_pls(X, ins.pop('Y'), permsamples=ins.pop('Yperm'),
                                  permindices=False, index_backend='reference',
                                  verbose=False, **ins)
    elif name.startswith('bpls'):
        out = pyls.behavioral_pls(X, ins.pop('Y'), index_backend='reference',
                                  verbose=False, **ins)
    else:
        out = pyls.meancentered_pls(X, index_backend='reference',
                                    verbose=False, **ins)
    if 'permsamples' in ref:
        assert np.array_equal(out.permres.permsamples, ref['permsamples'])
    keep = _non_null(ref['singvals'])
    np.testing.assert_allclose(out.permres.perm_singval[keep],
                               ref['perm_singval'][keep], rtol=1e-8,
                               atol=1e-11)
    for k in SPLIT_KEYS:
        np.testing.assert_allclose(out.splitres[k][keep], ref[k][keep],
                                   rtol=0, atol=ATOL, err_msg=k)
    for k in ('ucorr_pvals', 'vcorr_pvals'):
        assert np.array_equal(out.splitres[k][keep], ref[k][keep]), k


@pytest.mark.parametrize('groups,n_cond,T,cov', [
    ([20, 20], 2, 10, False),       # BASELINE config 2 layout (2 K = 80 rows)
    ([15], 1, 6, True),
    ([9, 11, 8], 2, 2, False),
    ([12], 3, 1, False),
    ([25, 25], 1, 50, False),       # K = 100: generic passes, global work space
])
def test_split_half_behavioral_matches_oracle(groups, n_cond, T, cov):
    """Engine level: permutations given as an index table, masks from the
    oracle's gen_splits, every permutation scored against its own
    decomposition; then pre-permuted Y matrices give the same numbers."""
    from pypyls_b200.engine import ResamplingEngine
    rs = np.random.RandomState(11)
    S, B, P, n_split = sum(groups) * n_cond, 900, 5, 4
    X, Y = rs.rand(S, B), rs.rand(S, T)
    Y[:, 0] += X[:, :20].mean(axis=1)
    spec = po._Spec('behavioral', groups, n_cond, covariance=cov, rotate=False)
    ps = po.gen_permsamp(groups, n_cond, P, seed=3)
    masks = np.stack([po.gen_splits(groups, n_cond, n_split, seed=i).T
                      for i in range(P)])
    want = [po.single_perm(spec, X, Y, ps[:, i], None, seed=i,
                           n_split=n_split) for i in range(P)]
    eng = ResamplingEngine('behavioral_cov' if cov else 'behavioral', S, B, T,
                           groups, n_cond)
    eng.set_data(X, Y)
    uc, vc = eng.split_half(masks, idx=ps)
    uc, vc = uc.cpu().numpy(), vc.cpu().numpy()
    for i in range(P):
        keep = _non_null(want[i][0])
        np.testing.assert_allclose(uc[i][keep], want[i][1][keep], rtol=0,
                                   atol=ATOL)
        np.testing.assert_allclose(vc[i][keep], want[i][2][keep], rtol=0,
                                   atol=ATOL)
    Yp = np.stack([Y[ps[:, i]] for i in range(P)])
    uc2, vc2 = eng.split_half(masks, Yperm=Yp)
    np.testing.assert_allclose(uc2.cpu().numpy(), uc, rtol=0, atol=1e-12)
    np.testing.assert_allclose(vc2.cpu().numpy(), vc, rtol=0, atol=1e-12)


@pytest.mark.parametrize('groups,n_cond,mc', [
    ([40, 40, 40, 40], 2, 0),       # BASELINE config 3 layout
    ([12, 9], 3, 1),
    ([12, 9], 3, 2),
    ([25, 30], 1, 1),
])
def test_split_half_meancentered_matches_oracle(groups, n_cond, mc):
    from pypyls_b200.engine import ResamplingEngine
    rs = np.random.RandomState(12)
    S, B, P, n_split = sum(groups) * n_cond, 600, 4, 5
    X = rs.rand(S, B)
    X[:groups[0]] += 0.2 * rs.rand(1, B)
    spec = po._Spec('meancentered', groups, n_cond, mean_centering=mc,
                    rotate=False)
    ps = po.gen_permsamp(groups, n_cond, P, seed=4)
    masks = np.stack([po.gen_splits(groups, n_cond, n_split, seed=i).T
                      for i in range(P)])
    eng = ResamplingEngine('meancentered', S, B, 1, groups, n_cond,
                           mean_centering=mc)
    eng.set_data(X)
    uc, vc = eng.split_half(masks, idx=ps)
    uc, vc = uc.cpu().numpy(), vc.cpu().numpy()
    for i in range(P):
        d, wu, wv = po.single_perm(spec, X, spec.dummy, ps[:, i], None,
                                   seed=i, n_split=n_split)
        keep = _non_null(d)
        assert keep.sum() >= 1
        np.testing.assert_allclose(uc[i][keep], wu[keep], rtol=0, atol=ATOL)
        np.testing.assert_allclose(vc[i][keep], wv[keep], rtol=0, atol=ATOL)
        assert np.all(uc[i][~keep] == 0) and np.all(vc[i][~keep] == 0)


def test_split_half_chunking_is_invisible():
    """A workspace too small for one permutation's halves makes the library
    loop over blocks of masks; the result must not change."""
    from pypyls_b200.engine import ResamplingEngine
    rs = np.random.RandomState(13)
    groups, n_cond, T, S, B, P, n_split = [10, 12], 2, 3, 44, 400, 6, 7
    X, Y = rs.rand(S, B), rs.rand(S, T)
    ps = po.gen_permsamp(groups, n_cond, P, seed=1)
    masks = np.stack([po.gen_splits(groups, n_cond, n_split, seed=i).T
                      for i in range(P)])
    res = []
    for ws in (None, 1 << 20):
        eng = ResamplingEngine('behavioral', S, B, T, groups, n_cond,
                               workspace_bytes=ws)
        eng.set_data(X, Y)
        uc, vc = eng.split_half(masks, idx=ps)
        res.append((uc.cpu().numpy(), vc.cpu().numpy()))
        eng.close()
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=0, atol=1e-12)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=0, atol=1e-12)


def test_split_half_properties_config2_size():
    """BASELINE config 2 shape (S=80, B=10000, T=10, groups [20, 20] x 2
    conditions; the pair of halves fills the 80-row limit).  Properties that do
    not need the oracle: exchanging the two halves of every mask leaves the
    correlations unchanged; values lie in [-1, 1]; plus one permutation
    against the oracle."""
    from pypyls_b200.engine import ResamplingEngine
    rs = np.random.RandomState(1234)
    groups, n_cond, T, S, B, P, n_split = [20, 20], 2, 10, 80, 10000, 12, 6
    X, Y = rs.rand(S, B), rs.rand(S, T)
    eng = ResamplingEngine('behavioral', S, B, T, groups, n_cond)
    eng.set_data(X, Y)
    ps, _ = eng.gen_perm_indices(99, P)
    masks, n_ex = eng.gen_split_masks(7, P, n_split)
    assert n_ex == 0
    uc, vc = eng.split_half(masks, idx=ps)
    uc2, vc2 = eng.split_half(1 - masks, idx=ps)
    uc, vc = uc.cpu().numpy(), vc.cpu().numpy()
    np.testing.assert_allclose(uc2.cpu().numpy(), uc, rtol=0, atol=1e-10)
    np.testing.assert_allclose(vc2.cpu().numpy(), vc, rtol=0, atol=1e-10)
    assert np.all(np.abs(uc) <= 1) and np.all(np.abs(vc) <= 1)
    assert np.all(np.isfinite(uc)) and np.all(np.isfinite(vc))
    spec = po._Spec('behavioral', groups, n_cond, rotate=False)
    perm0 = ps[0].cpu().numpy().astype(int)
    U, d, V = po.decompose(spec, X, Y[perm0], seed=0)
    di = np.linalg.inv(d)
    wu, wv = po.split_half(spec, X, Y[perm0], U @ di, V @ di, n_split,
                           splits=masks[0].cpu().numpy().T.astype(bool))
    np.testing.assert_allclose(uc[0], wu, rtol=0, atol=ATOL)
    np.testing.assert_allclose(vc[0], wv, rtol=0, atol=ATOL)


@pytest.mark.parametrize('groups,n_cond', [([10, 10], 2), ([20], 1),
                                           ([7, 9, 5], 3), ([4, 3], 1)])
def test_device_split_masks_obey_reference_invariants(groups, n_cond):
    """What gen_splits guarantees (pyls/base.py:162-229): per group ceil or
    floor of half the subjects, conditions follow their subject, no mask twice
    in a set; plus the counter-based generator's own contract (same seed ->
    same sets, a set does not depend on where the block starts)."""
    from pypyls_b200.engine import ResamplingEngine
    S, n_subj = sum(groups) * n_cond, sum(groups)
    n_sets, n_split = 9, 6 if groups == [4, 3] else 40
    eng = ResamplingEngine('meancentered' if len(groups) * n_cond > 1
                           else 'behavioral', S, 32, 1, groups, n_cond)
    masks, n_ex = eng.gen_split_masks(1234, n_sets, n_split)
    m = masks.cpu().numpy()
    assert m.shape == (n_sets, n_split, S) and n_ex == 0
    assert set(np.unique(m)) <= {0, 1}
    bounds = np.concatenate([[0], np.cumsum(groups)])
    for s in range(n_sets):
        assert len({row.tobytes() for row in m[s]}) == n_split
        for row in m[s]:
            for a, b in zip(bounds[:-1], bounds[1:]):
                blk = row[n_cond * a:n_cond * b].reshape(n_cond, b - a)
                assert np.all(blk == blk[0])
                assert blk[0].sum() in (int(np.floor((b - a) / 2)),
                                        int(np.ceil((b - a) / 2)))
    again, _ = eng.gen_split_masks(1234, n_sets, n_split)
    assert np.array_equal(again.cpu().numpy(), m)
    other, _ = eng.gen_split_masks(4321, n_sets, n_split)
    assert not np.array_equal(other.cpu().numpy(), m)
    tail, _ = eng.gen_split_masks(1234, n_sets - 4, n_split, first=4)
    assert np.array_equal(tail.cpu().numpy(), m[4:])
    # both halves are used about equally often by every subject
    frac = m[:, :, :].mean()
    assert 0.35 < frac < 0.65


def test_split_half_front_end_device_backend():
    """Default index_backend='device': tables and masks generated on the GPU;
    the results object carries the reference's splitres keys and the
    statistics are self-consistent."""
    import pypyls_b200 as pyls
    rs = np.random.RandomState(5)
    X, Y = rs.rand(40, 300), rs.rand(40, 3)
    Y[:, 0] += X[:, :30].mean(axis=1) * 3
    out = pyls.behavioral_pls(X, Y, groups=[10, 10], n_cond=2, n_perm=50,
                              n_boot=0, n_split=8, seed=3, verbose=False)
    sr = out.splitres
    L = 12
    for k in ('ucorr', 'vcorr', 'ucorr_pvals', 'vcorr_pvals', 'ucorr_lolim',
              'ucorr_uplim', 'vcorr_lolim', 'vcorr_uplim'):
        assert sr[k].shape == (L,) and np.all(np.isfinite(sr[k])), k
    assert np.all(sr.ucorr_lolim <= sr.ucorr_uplim)
    assert np.all((sr.ucorr_pvals > 0) & (sr.ucorr_pvals <= 1))
    again = pyls.behavioral_pls(X, Y, groups=[10, 10], n_cond=2, n_perm=50,
                                n_boot=0, n_split=8, seed=3, verbose=False)
    assert np.array_equal(again.splitres.ucorr, sr.ucorr)
    assert np.array_equal(again.splitres.vcorr_pvals, sr.vcorr_pvals)
