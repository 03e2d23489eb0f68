# -*- coding: utf-8 -*-
"""
Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (/root/reference, netneurolab/pypyls @ e0ff056) in the build
container.  The reference cannot travel to the GPU box, so its outputs are
committed as small .npz fixtures together with this script.

Shims (none touches arithmetic):
  1. a stub ``h5py`` module (tests/golden/_refshim) so ``import pyls`` works;
  2. ``permindices=True`` is always passed (HEAD treats the missing kwarg as
     False, pyls/base.py:628-639, 689-692);
  3. ``PLSRegression._single_perm`` is wrapped so it accepts the
     ``samples=/use_permind=`` call of pyls/base.py:646-648 (its signature at
     pyls/types/regression.py:329 is stale).
test_split=0 and n_split=0 except in the cross-validation and split-half
cases.

Run:  python tests/golden/make_golden.py
"""

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path.insert(0, os.path.join(HERE, '_refshim'))
sys.path.insert(0, REF)
warnings.filterwarnings('ignore')

import pyls  # noqa: E402
from pyls.types.regression import PLSRegression  # noqa: E402

_orig_sp = PLSRegression._single_perm


def _sp(self, X, Y, samples, use_permind=True, groups=None, original=None,
        seed=None):
    return _orig_sp(self, X, Y, inds=samples, groups=groups,
                    original=original, seed=seed)


PLSRegression._single_perm = _sp


def flat(res, kind):
    """PLSResults -> flat dict of arrays (only keys the hot path feeds)."""
    out = {}
    for k in ('x_weights', 'y_weights', 'x_scores', 'y_scores', 'y_loadings',
              'singvals', 'varexp'):
        v = res.get(k)
        if isinstance(v, np.ndarray):
            out[k] = v
    for k in ('pvals', 'permsamples', 'perm_singval'):
        v = res['permres'].get(k)
        if isinstance(v, np.ndarray):
            out[k] = v
    for k in ('x_weights_normed', 'x_weights_stderr', 'bootsamples',
              'y_loadings_boot', 'y_loadings_ci', 'contrast', 'contrast_boot',
              'contrast_ci'):
        v = res['bootres'].get(k)
        if isinstance(v, np.ndarray):
            out['boot_' + k if k in ('contrast',) else k] = v
    return out


def save(name, inputs, outputs):
    d = {'in_' + k: np.asarray(v) for k, v in inputs.items() if v is not None}
    d.update({'out_' + k: v for k, v in outputs.items()})
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **d)
    print(name, {k: v.shape for k, v in outputs.items()})


def linnerud():
    X = np.loadtxt(os.path.join(REF, 'data/linnerud_exercise.csv'),
                   delimiter=',', skiprows=1)[:, 1:]
    Y = np.loadtxt(os.path.join(REF, 'data/linnerud_physio.csv'),
                   delimiter=',', skiprows=1)[:, 1:]
    return X, Y


def behavioral_cases():
    X, Y = linnerud()
    kw = dict(n_perm=100, n_boot=100, seed=1234)
    r = pyls.behavioral_pls(X, Y, test_split=0, n_split=0, permindices=True,
                            verbose=False, **kw)
    save('bpls_linnerud', dict(X=X, Y=Y, groups=[20], n_cond=1, **kw),
         flat(r, 'b'))

    rs = np.random.RandomState(1234)
    X, Y = rs.rand(48, 60), rs.rand(48, 4)
    for tag, extra in (('rot', dict(rotate=True)),
                       ('norot', dict(rotate=False)),
                       ('cov', dict(covariance=True))):
        kw = dict(groups=[14, 10], n_cond=2, n_perm=40, n_boot=40, seed=4321,
                  **extra)
        r = pyls.behavioral_pls(X, Y, test_split=0, n_split=0,
                                permindices=True, verbose=False, **kw)
        save('bpls_2g2c_' + tag, dict(X=X, Y=Y, **kw), flat(r, 'b'))

    # one group / one condition, wider Y (K = T = 12), user-given resamples
    X, Y = rs.rand(30, 200), rs.rand(30, 12)
    ps = pyls.base.gen_permsamp([30], 1, 25, seed=7, verbose=False)
    bs = pyls.base.gen_bootsamp([30], 1, 25, seed=8, verbose=False)
    kw = dict(n_perm=25, n_boot=25, seed=99)
    r = pyls.behavioral_pls(X, Y, test_split=0, n_split=0, permindices=True,
                            verbose=False, permsamples=ps, bootsamples=bs,
                            **kw)
    save('bpls_1g1c_given', dict(X=X, Y=Y, groups=[30], n_cond=1,
                                 permsamples=ps, bootsamples=bs, **kw),
         flat(r, 'b'))


def prepermuted_case():
    """`permsamples` as a (P, S, T) stack of pre-permuted Y matrices with
    permindices=False (pyls/base.py:636-639, 689-692): the spatial-null use."""
    rs = np.random.RandomState(4242)
    X, Y = rs.rand(36, 90), rs.rand(36, 5)
    Y[:, 0] += X[:, :8].mean(axis=1)
    groups, n_cond, P = [10, 8], 2, 30
    Yp = np.stack([Y[rs.permutation(36)] + 0.05 * rs.rand(36, 5)
                   for _ in range(P)])
    for tag, rot in (('rot', True), ('norot', False)):
        kw = dict(groups=groups, n_cond=n_cond, n_perm=P, n_boot=0, seed=11,
                  rotate=rot)
        r = pyls.behavioral_pls(X, Y, test_split=0, n_split=0,
                                permsamples=Yp, permindices=False,
                                verbose=False, **kw)
        out = flat(r, 'b')
        out.pop('permsamples', None)
        save('bpls_prepermuted_' + tag, dict(X=X, Y=Y, Yperm=Yp, **kw), out)


def crossval_case():
    """Cross-validation (test_split train / test splits, BehavioralPLS.crossval,
    pyls/types/behavioral.py:82-170)."""
    rs = np.random.RandomState(777)
    X, Y = rs.rand(48, 60), rs.rand(48, 4)
    Y[:, :2] += X[:, :12] @ rs.rand(12, 2) * 0.4
    for tag, extra in (('corr', {}), ('cov', dict(covariance=True))):
        kw = dict(groups=[14, 10], n_cond=2, n_perm=10, n_boot=10, seed=31,
                  test_split=15, test_size=0.25, **extra)
        r = pyls.behavioral_pls(X, Y, n_split=0, permindices=True,
                                verbose=False, **kw)
        out = flat(r, 'b')
        out['pearson_r'] = r['cvres']['pearson_r']
        out['r_squared'] = r['cvres']['r_squared']
        save('bpls_crossval_' + tag, dict(X=X, Y=Y, **kw), out)


def splithalf_cases():
    """Split-half resampling inside the permutation loop (n_split,
    pyls/base.py:373-397, 704-708, 714-770)."""
    keys = ('ucorr', 'vcorr', 'ucorr_pvals', 'vcorr_pvals', 'ucorr_lolim',
            'ucorr_uplim', 'vcorr_lolim', 'vcorr_uplim')
    rs = np.random.RandomState(99)
    X, Y = rs.rand(48, 60), rs.rand(48, 4)
    Y[:, :2] += X[:, :12] @ rs.rand(12, 2) * 0.4
    for tag, extra in (('rot', dict(rotate=True)),
                       ('cov_norot', dict(rotate=False, covariance=True))):
        kw = dict(groups=[14, 10], n_cond=2, n_perm=20, n_boot=0, seed=5,
                  n_split=6, **extra)
        r = pyls.behavioral_pls(X, Y, test_split=0, permindices=True,
                                verbose=False, **kw)
        out = flat(r, 'b')
        out.update({k: np.asarray(r['splitres'][k]) for k in keys})
        save('bpls_split_' + tag, dict(X=X, Y=Y, **kw), out)
    X = rs.rand(48, 50)
    X[:16] += 0.3 * rs.rand(1, 50)
    for mc in (0, 1, 2):
        kw = dict(groups=[8, 10, 6], n_cond=2, mean_centering=mc, n_perm=20,
                  n_boot=0, seed=5, n_split=6)
        r = pyls.meancentered_pls(X, permindices=True, verbose=False, **kw)
        out = flat(r, 'm')
        out.update({k: np.asarray(r['splitres'][k]) for k in keys})
        save('mpls_split_mc%d' % mc, dict(X=X, **kw), out)
    splithalf_more_cases(keys)


def splithalf_more_cases(keys):
    """Split-half with pre-permuted Y matrices (permindices=False: the masks
    split X and the handed-in Y, pyls/base.py:689-692, 704-708) and with
    uneven groups x three conditions."""
    rs = np.random.RandomState(4243)
    X, Y = rs.rand(36, 90), rs.rand(36, 5)
    Y[:, 0] += X[:, :8].mean(axis=1)
    groups, n_cond, P = [10, 8], 2, 14
    Yp = np.stack([Y[rs.permutation(36)] + 0.05 * rs.rand(36, 5)
                   for _ in range(P)])
    kw = dict(groups=groups, n_cond=n_cond, n_perm=P, n_boot=0, seed=11,
              n_split=5, rotate=True)
    r = pyls.behavioral_pls(X, Y, test_split=0, permsamples=Yp,
                            permindices=False, verbose=False, **kw)
    out = flat(r, 'b')
    out.pop('permsamples', None)
    out.update({k: np.asarray(r['splitres'][k]) for k in keys})
    save('bpls_split_prepermuted', dict(X=X, Y=Y, Yperm=Yp, **kw), out)

    X, Y = rs.rand(57, 70), rs.rand(57, 2)
    kw = dict(groups=[7, 5, 7], n_cond=3, n_perm=12, n_boot=0, seed=3,
              n_split=4)
    r = pyls.behavioral_pls(X, Y, test_split=0, permindices=True,
                            verbose=False, **kw)
    out = flat(r, 'b')
    out.update({k: np.asarray(r['splitres'][k]) for k in keys})
    save('bpls_split_3g3c', dict(X=X, Y=Y, **kw), out)


def meancentered_cases():
    rs = np.random.RandomState(1234)
    X = rs.rand(48, 50)
    X[:16] += 0.3 * rs.rand(1, 50)          # a little group structure
    for mc in (0, 1, 2):
        for tag, rot in (('rot', True), ('norot', False)):
            if mc != 0 and not rot:
                continue
            kw = dict(groups=[8, 10, 6], n_cond=2, mean_centering=mc,
                      n_perm=40, n_boot=40, seed=2468, rotate=rot)
            r = pyls.meancentered_pls(X, n_split=0, permindices=True,
                                      verbose=False, **kw)
            save('mpls_3g2c_mc%d_%s' % (mc, tag), dict(X=X, **kw),
                 flat(r, 'm'))


def regression_cases():
    rs = np.random.RandomState(1234)
    for tag, (S, B, T, L) in (('t12', (40, 60, 12, 4)), ('t5', (36, 80, 5, 3))):
        X, Y = rs.rand(S, B), rs.rand(S, T)
        Y[:, :2] += X[:, :10] @ rs.rand(10, 2) * 0.3
        kw = dict(n_components=L, n_perm=30, n_boot=30, seed=1357)
        r = pyls.pls_regression(X.copy(), Y.copy(), permindices=True,
                                verbose=False, **kw)
        save('plsr_' + tag, dict(X=X, Y=Y, **kw), flat(r, 'r'))


def regression_missing_rows_case():
    """Rows of X / Y that are missing altogether (all NaN): masked by get_mask
    (pyls/types/regression.py:48-53) in the original decomposition and, on the
    resampled matrices, in every permutation and bootstrap."""
    rs = np.random.RandomState(3)
    X, Y = rs.rand(40, 60), rs.rand(40, 5)
    Y[:, :2] += X[:, :10] @ rs.rand(10, 2) * 0.3
    X[17, :] = np.nan
    Y[25, :] = np.nan
    X[31, :] = np.nan
    kw = dict(n_components=3, n_perm=12, n_boot=12, seed=5)
    r = pyls.pls_regression(X.copy(), Y.copy(), permindices=True,
                            verbose=False, **kw)
    save('plsr_missing_rows', dict(X=X, Y=Y, **kw), flat(r, 'r'))


def regression_3d_cases():
    """Three-dimensional Y (S, T, C) with `aggfunc` (pyls/types/regression.py:
    207-235, 308-310): every bootstrap aggregates a bootstrap sample of the
    third axis into its own behaviour matrix.  The reference's own table
    generation for this case does not run under NumPy 2 (inhomogeneous
    np.array, regression.py:215), so the (2, n_boot) object table is built
    here from its two gen_bootsamp draws and handed in, which is the
    reference's documented alternative (regression.py:217-228)."""
    rs = np.random.RandomState(77)
    S, B, T, C, n_boot = 40, 120, 5, 6, 12
    X, Y = rs.rand(S, B), rs.rand(S, T, C)
    s = pyls.base.gen_bootsamp([S], 1, n_boot, seed=5, verbose=False)
    c = pyls.base.gen_bootsamp([C], 1, n_boot, seed=6, verbose=False)
    bs = np.empty((2, n_boot), dtype=object)
    for i in range(n_boot):
        bs[0, i], bs[1, i] = s[:, i], c[:, i]
    for agg in ('mean', 'median'):
        kw = dict(n_components=4, n_perm=8, n_boot=n_boot, seed=11)
        r = pyls.pls_regression(X.copy(), Y.copy(), permindices=True,
                                verbose=False, bootsamples=bs, aggfunc=agg,
                                **kw)
        out = flat(r, 'r')
        out.pop('bootsamples', None)          # object array: kept as two tables
        save('plsr_3d_' + agg, dict(X=X, Y=Y, boot_rows=s, boot_third=c, **kw),
             out)


def index_cases():
    out = {}
    for n, (groups, n_cond, seed, cnt) in enumerate((
            ([10, 10], 2, 1234, 10), ([20], 1, 1, 50), ([7, 9, 5], 3, 42, 30),
            ([12], 2, 5, 20), ([4, 3], 1, 3, 12))):
        out['perm%d' % n] = pyls.base.gen_permsamp(groups, n_cond, cnt,
                                                   seed=seed, verbose=False)
        out['boot%d' % n] = pyls.base.gen_bootsamp(groups, n_cond, cnt,
                                                   seed=seed, verbose=False)
        out['spec%d' % n] = np.array(groups + [n_cond, seed, cnt])
    np.savez_compressed(os.path.join(HERE, 'index_tables.npz'), **out)
    print('index_tables', len(out))


def matlab_cases():
    """In-tree Matlab PLS-toolbox fixtures (pyls/tests/data/*.mat): inputs,
    the Matlab-generated resampling tables and Matlab's results, plus what the
    reference itself produces from them (pyls/tests/matlab.py:202-285)."""
    for name in ('bpls_onegroup_onecond_nosplit',
                 'mpls_multigroup_onecond_nosplit'):
        m = pyls.matlab.import_matlab_result(
            os.path.join(REF, 'pyls/tests/data', name + '.mat'))
        inp = dict(m['inputs'])
        fcn = (pyls.behavioral_pls if inp['method'] == 3
               else pyls.meancentered_pls)
        keep = {k: inp[k] for k in ('X', 'Y', 'groups', 'n_cond', 'n_perm',
                                    'n_boot', 'mean_centering', 'permsamples',
                                    'bootsamples', 'rotate', 'ci')
                if inp.get(k) is not None}
        if fcn is pyls.meancentered_pls:
            keep.pop('Y', None)
        run = dict(keep, seed=1234, verbose=False, n_split=0,
                   permindices=True)
        if fcn is pyls.behavioral_pls:
            run['test_split'] = 0
        py = fcn(**run)
        # the full bootstrap distributions are large; their CIs are kept
        drop = ('permsamples', 'bootsamples', 'y_loadings_boot',
                'contrast_boot')
        out = {'py_' + k: v for k, v in flat(py, '').items()
               if k not in drop}
        out.update({'ml_' + k: v for k, v in flat(m, '').items()
                    if k not in drop})
        save('matlab_' + name, keep, out)


if __name__ == '__main__':
    if sys.argv[1:] == ['splithalf']:
        splithalf_cases()
        sys.exit(0)
    if sys.argv[1:] == ['splitmore']:
        splithalf_more_cases(('ucorr', 'vcorr', 'ucorr_pvals', 'vcorr_pvals',
                              'ucorr_lolim', 'ucorr_uplim', 'vcorr_lolim',
                              'vcorr_uplim'))
        sys.exit(0)
    if sys.argv[1:] == ['missing']:
        regression_missing_rows_case()
        sys.exit(0)
    if sys.argv[1:] == ['regression3d']:
        regression_3d_cases()
        sys.exit(0)
    behavioral_cases()
    splithalf_cases()
    prepermuted_case()
    crossval_case()
    meancentered_cases()
    regression_cases()
    regression_missing_rows_case()
    regression_3d_cases()
    index_cases()
    matlab_cases()
