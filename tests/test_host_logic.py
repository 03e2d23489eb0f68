# -*- coding: utf-8 -*-
"""
CPU tests of the host side: resampling-table replay, result containers, and
that the C-ABI library loads and exports every symbol include/plsb200.h
declares (no compute calls -- there is no GPU here).
"""

import os
import re
import warnings

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden
from oracle import pls_oracle as po


def test_index_tables_bit_exact_with_reference():
    from pypyls_b200.resample import gen_bootsamp, gen_permsamp
    z = np.load(os.path.join(GOLDEN, 'index_tables.npz'))
    n = 0
    while 'spec%d' % n in z.files:
        spec = [int(v) for v in z['spec%d' % n]]
        groups, (n_cond, seed, cnt) = spec[:-3], spec[-3:]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            assert np.array_equal(gen_permsamp(groups, n_cond, cnt, seed=seed),
                                  z['perm%d' % n])
            assert np.array_equal(gen_bootsamp(groups, n_cond, cnt, seed=seed),
                                  z['boot%d' % n])
        n += 1
    assert n == 5


@pytest.mark.parametrize('groups,n_cond', [([6, 5], 2), ([9], 3), ([3, 3, 4], 1)])
def test_index_tables_match_oracle_stream(groups, n_cond):
    from pypyls_b200.resample import gen_bootsamp, gen_permsamp
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        a = gen_permsamp(groups, n_cond, 40, seed=np.random.RandomState(5))
        b = po.gen_permsamp(groups, n_cond, 40, seed=np.random.RandomState(5))
        assert np.array_equal(a, b)
        a = gen_bootsamp(groups, n_cond, 40, seed=np.random.RandomState(6))
        b = po.gen_bootsamp(groups, n_cond, 40, seed=np.random.RandomState(6))
        assert np.array_equal(a, b)


def test_duplicate_warning_on_tiny_layout():
    """pyls/tests/test_base.py: tiny inputs exhaust the distinct resamples."""
    from pypyls_b200.resample import gen_bootsamp, gen_permsamp
    with pytest.warns(UserWarning, match='Duplicate permutations'):
        gen_permsamp([2], 1, 5, seed=1)
    with pytest.warns(UserWarning, match='Duplicate bootstraps'):
        gen_bootsamp([3], 1, 30, seed=1)


@pytest.mark.parametrize('name', ['bpls_2g2c_rot', 'mpls_3g2c_mc0_rot'])
def test_seed_replay_reproduces_reference_tables(name):
    """The analysis' RandomState is consumed by the original decomposition
    (a (K, K+10) normal draw inside sklearn's randomized_svd) before the
    tables are made (pyls/base.py:362-368, 466-470)."""
    from pypyls_b200.resample import (check_random_state, gen_bootsamp,
                                      gen_permsamp)
    ins, ref = load_golden(name)
    J = len(ins['groups']) * ins['n_cond']
    K = J * ins['Y'].shape[1] if 'Y' in ins else J
    rs = check_random_state(ins['seed'])
    rs.normal(size=(K, K + 10))
    perm = gen_permsamp(ins['groups'], ins['n_cond'], ins['n_perm'], seed=rs)
    boot = gen_bootsamp(ins['groups'], ins['n_cond'], ins['n_boot'], seed=rs)
    assert np.array_equal(perm, ref['permsamples'])
    assert np.array_equal(boot, ref['bootsamples'])


def test_shard_range_partitions():
    from pypyls_b200.resample import shard_range
    for n in (0, 1, 7, 5000, 10001):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0
            for (f0, c0), (f1, _) in zip(blocks[:-1], blocks[1:]):
                assert f0 + c0 == f1
            assert blocks[-1][0] + blocks[-1][1] == n
            sizes = [c for _, c in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_structures_behave_like_reference():
    """pyls/tests/test_structures.py: whitelist, n_split/test_split/n_proc
    normalisation, test_size validation, tolerant equality."""
    from pypyls_b200.structures import PLSInputs, PLSResults
    inp = PLSInputs(X=1, n_split=0, test_split=0, n_proc='max', bogus=3)
    assert 'bogus' not in inp and inp.n_split is None
    assert inp.test_split is None and inp.n_proc == os.cpu_count()
    assert PLSInputs(n_proc=-2).n_proc == os.cpu_count() - 1
    with pytest.raises(ValueError):
        PLSInputs(test_size=1)
    with pytest.raises(ValueError):
        PLSInputs(test_size=-0.5)
    a = PLSResults(x_weights=np.ones(3), n_perm=5)
    b = PLSResults(x_weights=np.ones(3) + 1e-7, n_perm=5)
    c = PLSResults(x_weights=np.ones(3) + 1e-5, n_perm=5)
    assert a == b and a != c
    assert a.inputs.n_perm == 5 and 'x_weights' in str(a)
    a.not_allowed = 4
    assert 'not_allowed' not in a


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'plsb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(plsb_[a-z_0-9]+)\s*\(', text)))


def test_cabi_exports_every_declared_symbol():
    """Builds (if needed) and loads libplsb200.so; every function declared in
    include/plsb200.h must be exported and bound by the ctypes layer."""
    from pypyls_b200 import _build, _cabi
    _build.build()
    lib = _cabi.load()
    names = _header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
        assert name in _cabi.PROTOTYPES, name
    assert set(_cabi.PROTOTYPES) == set(names)
    assert lib.plsb_version() >= 100


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from pypyls_b200.engine import ResamplingEngine
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ResamplingEngine('behavioral', 8, 16, 2, [8])
    import pypyls_b200 as pyls
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pyls.behavioral_pls(np.random.rand(8, 16), np.random.rand(8, 2),
                            n_perm=2, n_boot=2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'pypyls_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text,
                                     flags=re.M), f
                assert 'pls_oracle' not in text, f


def test_regression_host_helpers_match_reference_semantics():
    """gaussian_tables reproduces sklearn's draw for an integer seed
    (RandomState(i).normal(size=(T, 11))); resid_yscores equals the oracle's
    restatement of pyls/types/regression.py:9-45."""
    from pypyls_b200.types.regression import gaussian_tables, resid_yscores
    tab = gaussian_tables([0, 7, 123456], 13)
    for n, i in enumerate([0, 7, 123456]):
        assert np.array_equal(tab[n],
                              np.random.RandomState(i).normal(size=(13, 11)))
    rs = np.random.RandomState(3)
    xs, ys = rs.rand(30, 5), rs.rand(30, 5)
    np.testing.assert_allclose(resid_yscores(xs, ys),
                               po.resid_yscores(xs, ys), rtol=1e-13)


def test_regression_front_end_argument_errors():
    """pyls/tests/types/test_regression.py: n_components bound, aggfunc and
    the paired bootstrap tables of a three-dimensional Y, checked before the
    GPU is touched."""
    import pypyls_b200 as pyls
    rs = np.random.RandomState(1)
    X, Y = rs.rand(20, 30), rs.rand(20, 4)
    with pytest.raises(ValueError, match='n_components'):
        pyls.pls_regression(X, Y, n_components=25, n_perm=0, n_boot=0)
    Y3 = rs.rand(20, 4, 3)
    with pytest.raises(ValueError, match='aggfunc'):
        pyls.pls_regression(X, Y3, n_perm=0, n_boot=0, aggfunc='mode')
    with pytest.raises(ValueError, match='bootsamples'):
        pyls.pls_regression(X, Y3, n_perm=0, n_boot=4,
                            bootsamples=np.zeros((20, 4), dtype=int))
    # rows that are missing altogether are masked (on the GPU); any other NaN
    # fails like it does inside the reference's randomized_svd
    Xn = X.copy()
    Xn[3, 5] = np.nan
    with pytest.raises(ValueError, match='NaN'):
        pyls.pls_regression(Xn, Y, n_perm=0, n_boot=0)


@pytest.mark.parametrize('groups,n_cond,test_size', [([20, 20], 2, 0.25),
                                                     ([7, 9, 5], 3, 0.5),
                                                     ([30], 1, 0.25)])
def test_gen_splits_replays_the_reference_stream(groups, n_cond, test_size):
    """The product's gen_splits draws what the oracle's restatement of
    pyls/base.py:162-229 draws (the oracle is pinned against the reference by
    the cross-validation golden vectors)."""
    from oracle import pls_oracle as po
    from pypyls_b200.resample import gen_splits
    a = gen_splits(groups, n_cond, 40, seed=5, test_size=test_size)
    b = po.gen_splits(groups, n_cond, 40, seed=5, test_size=test_size)
    assert a.dtype == bool and np.array_equal(a, b)
    assert len(set(map(bytes, a.T))) == 40          # no duplicate splits
    # conditions follow their subject
    n_subj = sum(groups)
    assert a.sum() % n_cond == 0 and a.shape[0] == n_subj * n_cond


def test_results_io_round_trip(tmp_path):
    """save_results / load_results keep the reference's HDF5 layout
    (pyls/io.py:12-122; pyls/tests/test_io.py): arrays as datasets, scalars as
    attributes, None as the string 'None'.  With h5py installed this writes and
    reads real HDF5; without it (this image) the same calls run against a
    stand-in with the h5py API (tests/_fake_h5py.py), which still checks what
    becomes a group, a dataset and an attribute."""
    try:
        import h5py
    except ImportError:
        import sys
        import _fake_h5py as h5py
        sys.modules['h5py'] = h5py
    import pypyls_b200 as pyls
    from pypyls_b200.structures import PLSResults
    rs = np.random.RandomState(0)
    res = PLSResults(x_weights=rs.rand(5, 2), singvals=rs.rand(2),
                     inputs=dict(X=rs.rand(4, 5), n_perm=3, seed=None,
                                 groups=[4]),
                     permres=dict(pvals=rs.rand(2)))
    fname = pyls.save_results(tmp_path / 'res', res)
    assert fname.endswith('.hdf5') and h5py.is_hdf5(fname)
    back = pyls.load_results(fname)
    np.testing.assert_array_equal(back.x_weights, res.x_weights)
    np.testing.assert_array_equal(back.permres.pvals, res.permres.pvals)
    assert back.inputs.n_perm == 3 and back.inputs.seed is None
    with h5py.File(fname, 'r') as f:
        assert f['/results/inputs'].attrs['seed'] == 'None'
        assert isinstance(f['/results/permres/pvals'], h5py.Dataset)
    with pytest.raises(TypeError):
        (tmp_path / 'junk.hdf5').write_text('not hdf5')
        pyls.load_results(tmp_path / 'junk.hdf5')
    if h5py.__name__ == '_fake_h5py':
        import sys
        del sys.modules['h5py']


def test_bootsamples_must_stay_inside_their_cells():
    """The engine's cell-grouped operands assume what gen_bootsamp guarantees
    (pyls/base.py:134-143); a table that breaks it is refused, not mangled."""
    from pypyls_b200.resample import (cell_of_rows, check_bootsamples,
                                      gen_bootsamp, gen_permsamp)
    groups, n_cond = [5, 4], 2
    assert cell_of_rows(groups, n_cond).tolist() == \
        [0] * 5 + [1] * 5 + [2] * 4 + [3] * 4
    check_bootsamples(gen_bootsamp(groups, n_cond, 30, seed=1, verbose=False),
                      groups, n_cond)
    with pytest.raises(ValueError, match='cells'):
        check_bootsamples(gen_permsamp(groups, n_cond, 5, seed=1,
                                       verbose=False), groups, n_cond)
