# -*- coding: utf-8 -*-
"""
GPU unit tests of the C-ABI primitives against NumPy / the CPU oracle.
Every call goes through libplsb200.so (pypyls_b200.engine is a ctypes shim).
"""

import numpy as np
import pytest

from oracle import pls_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device')
    return torch


def make_engine(mode, S, B, T, groups, n_cond=1, mean_centering=0, **kw):
    from pypyls_b200.engine import ResamplingEngine
    return ResamplingEngine(mode, S, B, T, groups, n_cond,
                            mean_centering=mean_centering, **kw)


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


# ---------------------------------------------------------------------------
@pytest.mark.parametrize('M,N,Kd', [(1, 1, 1), (7, 5, 3), (128, 128, 16),
                                    (300, 1000, 80), (129, 257, 33),
                                    (1000, 130, 200)])
def test_dgemm_matches_numpy(torch_cuda, M, N, Kd):
    """The FP64 DMMA kernel against NumPy."""
    rs = np.random.RandomState(M * 7 + N)
    A, X = rs.randn(M, Kd), rs.randn(Kd, N)
    eng = make_engine('behavioral', 4, 8, 1, [4], gemm_backend='dmma')
    C = eng.dgemm(A, X).cpu().numpy()
    np.testing.assert_allclose(C, A @ X, rtol=1e-12, atol=1e-12)


# tolerance per number of int8 digit planes, relative to |a_row| |x_col| per entry
SLICE_TOL = {5: 2e-10, 6: 1e-12, 7: 8e-15}


@pytest.mark.parametrize('slices', [5, 6, 7])
@pytest.mark.parametrize('M,N,Kd', [(1, 1, 1), (7, 5, 3), (128, 128, 32),
                                    (300, 1000, 80), (129, 257, 33),
                                    (1000, 130, 200), (513, 300, 224)])
def test_slice_gemm_matches_numpy(torch_cuda, M, N, Kd, slices):
    """The int8 slice GEMM on the tcgen05 tensor cores (csrc/gemm_i8.cu) against
    NumPy: entry-wise error bounded relative to the norms of the row of A and the
    column of X it contracts (the bound of a floating-point dot product)."""
    rs = np.random.RandomState(M * 7 + N + slices)
    # rows of A / columns of X on very different scales: every row and column
    # carries its own power of two
    A = rs.randn(M, Kd) * 10.0 ** rs.randint(-6, 7, size=(M, 1))
    X = rs.randn(Kd, N) * 10.0 ** rs.randint(-6, 7, size=(1, N))
    eng = make_engine('behavioral', 4, 8, 1, [4], gemm_backend='auto',
                      gemm_slices=slices)
    C = eng.dgemm(A, X).cpu().numpy()
    norm = np.sqrt((A ** 2).sum(1))[:, None] * np.sqrt((X ** 2).sum(0))[None, :]
    err = np.abs(C - A @ X) / norm
    assert err.max() < SLICE_TOL[slices], err.max()


@pytest.mark.parametrize('slices', [5, 6, 7])
@pytest.mark.parametrize('M,N,Kd', [(7, 5, 3), (129, 257, 33), (300, 1000, 200)])
def test_slice_gemm_bit_equal_to_restatement(torch_cuda, M, N, Kd, slices):
    """Integer work is exact and the FP64 recombination has a fixed order: the CUDA
    kernel equals the NumPy restatement of its arithmetic (oracle/slice_gemm.py)
    bit for bit."""
    from oracle.slice_gemm import slice_gemm
    rs = np.random.RandomState(M + N + slices)
    A = rs.randn(M, Kd) * 10.0 ** rs.randint(-3, 4, size=(M, 1))
    X = rs.randn(Kd, N) * 10.0 ** rs.randint(-3, 4, size=(1, N))
    A[M // 2] = 0.0
    eng = make_engine('behavioral', 4, 8, 1, [4], gemm_backend='auto',
                      gemm_slices=slices)
    C = eng.dgemm(A, X).cpu().numpy()
    assert np.array_equal(C, slice_gemm(A, X, slices))


@pytest.mark.parametrize('amax', [1, 7, 63, 64, 300, 70000])
def test_slice_gemm_skips_empty_planes(torch_cuda, amax):
    """Operands of small integers (the multiplicity rows of the bootstrap column
    statistics) fill only their leading digit planes; the kernel skips the planes
    that are all zero.  Same bits as the restatement, and exact while the product
    is representable."""
    from oracle.slice_gemm import slice_gemm
    rs = np.random.RandomState(amax)
    A = rs.randint(0, amax + 1, size=(260, 200)).astype(float)
    A[:, 17] = amax
    X = rs.randn(200, 400)
    eng = make_engine('behavioral', 4, 8, 1, [4], gemm_backend='auto')
    C = eng.dgemm(A, X).cpu().numpy()
    assert np.array_equal(C, slice_gemm(A, X, 6))
    Xi = rs.randint(-100, 101, size=(200, 400)).astype(float)
    assert np.array_equal(eng.dgemm(A, Xi).cpu().numpy(), A @ Xi)


def test_slice_gemm_special_values(torch_cuda):
    """Zero rows / columns stay exactly zero, NaN and Inf poison exactly the rows /
    columns a product would, contractions beyond 224 rows take the DMMA kernel."""
    rs = np.random.RandomState(11)
    eng = make_engine('behavioral', 4, 8, 1, [4], gemm_backend='auto')
    A, X = rs.randn(200, 100), rs.randn(100, 300)
    A[5] = 0.0
    X[:, 7] = 0.0
    A[3, 5] = np.nan
    X[7, 9] = np.inf
    C = eng.dgemm(A, X).cpu().numpy()
    assert np.all(C[5, np.arange(300) != 9] == 0.0)
    assert np.all(C[np.arange(200) != 3, 7] == 0.0)
    assert np.isnan(C[3]).all() and not np.isfinite(C[:, 9]).any()
    rest = np.delete(np.delete(C, 3, 0), 9, 1)
    ref = np.delete(np.delete(A, 3, 0) @ np.delete(X, 9, 1), [], 0)
    assert np.isfinite(rest).all()
    assert rel_err(rest, ref) < 1e-11
    A, X = rs.randn(64, 300), rs.randn(300, 200)      # Kd > 224
    np.testing.assert_allclose(eng.dgemm(A, X).cpu().numpy(), A @ X,
                               rtol=1e-12, atol=1e-12)
    # tiny and huge magnitudes: the powers of two carry them
    A, X = rs.randn(130, 64) * 1e-150, rs.randn(64, 140) * 1e140
    assert rel_err(eng.dgemm(A, X).cpu().numpy(), A @ X) < 1e-11


def test_dgemm_linearity_large(torch_cuda):
    """Size-independent property at a benchmark-sized contraction:
    (aA1 + bA2) X == a(A1 X) + b(A2 X)."""
    rs = np.random.RandomState(5)
    A1, A2, X = rs.randn(2000, 80), rs.randn(2000, 80), rs.randn(80, 10000)
    for backend, tol in (('dmma', 1e-12), ('auto', 1e-11)):
        eng = make_engine('behavioral', 4, 8, 1, [4], gemm_backend=backend)
        lhs = eng.dgemm(2.0 * A1 - 3.0 * A2, X)
        rhs = 2.0 * eng.dgemm(A1, X) - 3.0 * eng.dgemm(A2, X)
        assert rel_err(lhs.cpu().numpy(), rhs.cpu().numpy()) < tol


# ---------------------------------------------------------------------------
@pytest.mark.parametrize('K', [1, 2, 3, 5, 10, 25, 40, 64, 80])
def test_small_decomp_matches_lapack(torch_cuda, K):
    """M = V Q reproduces compute.procrustes on top of an exact SVD:
    R^T M == U d P^T N^T (pyls/compute.py:240-264)."""
    rs = np.random.RandomState(K)
    n, B = 6, max(3 * K, 50)
    eng = make_engine('behavioral', 4, 8, 1, [4])
    Rs = rs.randn(n, K, B)
    Uo = np.linalg.qr(rs.randn(B, K))[0]
    G = np.einsum('rkb,rjb->rkj', Rs, Rs)
    H = np.einsum('rkb,bl->rkl', Rs, Uo)
    M, lam = eng.small_decomp(G, H)
    M, lam = M.cpu().numpy(), lam.cpu().numpy()
    for r in range(n):
        U, d, Vt = np.linalg.svd(Rs[r].T, full_matrices=False)
        np.testing.assert_allclose(np.sqrt(lam[r]), d, rtol=1e-10)
        N, _, P = np.linalg.svd(Uo.T @ U)
        want = U @ np.diag(d) @ (P.T @ N.T)
        got = Rs[r].T @ M[r]
        assert rel_err(got, want) < 1e-9, (K, r, rel_err(got, want))


@pytest.mark.parametrize('K', [1, 2, 3, 7, 8, 9, 10, 16, 17, 24, 25, 33, 40, 41,
                               48, 49, 57, 64, 65, 73, 80])
@pytest.mark.parametrize('B', [50, 128, 777])
def test_gram_proj_and_accum_u_match_numpy(torch_cuda, K, B):
    """The two streaming DMMA contractions around the small decomposition, for
    every fragment count (K / 8) and for column counts that are not multiples
    of the stage width."""
    rs = np.random.RandomState(1000 * K + B)
    n = 5
    eng = make_engine('behavioral', 4, 8, 1, [4])
    Rs = rs.randn(n, K, B)
    Uo = rs.randn(B, K)
    G, H = eng.gram_proj(Rs, Uo)
    G, H = G.cpu().numpy(), H.cpu().numpy()
    wantG = np.einsum('rkb,rjb->rkj', Rs, Rs)
    assert rel_err(G, wantG) < 1e-13
    assert np.array_equal(G, G.transpose(0, 2, 1))       # exactly symmetric
    assert rel_err(H, np.einsum('rkb,bl->rkl', Rs, Uo)) < 1e-13
    G2, H2 = eng.gram_proj(Rs)                            # Gram matrix only
    assert H2 is None and rel_err(G2.cpu().numpy(), wantG) < 1e-13
    M = rs.randn(n, K, K)
    us, uq = eng.accum_u(Rs, M)
    U = np.einsum('rkb,rkl->rbl', Rs, M)
    assert rel_err(us.cpu().numpy(), U.sum(0)) < 1e-12
    assert rel_err(uq.cpu().numpy(), (U ** 2).sum(0)) < 1e-12


def test_small_decomp_rank_deficient(torch_cuda):
    """A numerically null direction (mean-centred PLS always has one) must
    not poison the other latent variables."""
    rs = np.random.RandomState(3)
    K, B, n = 6, 200, 4
    eng = make_engine('behavioral', 4, 8, 1, [4])
    Rs = rs.randn(n, K, B)
    Rs -= Rs.mean(axis=1, keepdims=True)          # rows sum to zero -> rank K-1
    Uo, do, _ = np.linalg.svd(Rs[0].T, full_matrices=False)
    G = np.einsum('rkb,rjb->rkj', Rs, Rs)
    H = np.einsum('rkb,bl->rkl', Rs, Uo)
    M, lam = eng.small_decomp(G, H, d_orig=do)
    M = M.cpu().numpy()
    assert np.all(np.isfinite(M))
    assert np.abs(np.einsum('rkb,rkl->rbl', Rs, M)[..., -1]).max() < 1e-9
    for r in range(n):
        U, d, Vt = np.linalg.svd(Rs[r].T, full_matrices=False)
        U, d = U[:, :K - 1], d[:K - 1]
        N, _, P = np.linalg.svd(Uo[:, :K - 1].T @ U)
        want = U @ np.diag(d) @ (P.T @ N.T)
        got = (Rs[r].T @ M[r])[:, :K - 1]
        assert rel_err(got, want) < 1e-6


# ---------------------------------------------------------------------------
@pytest.mark.parametrize('count', [1, 2, 3, 100, 511, 512, 513, 1000, 5000,
                                   20001, 40000, 70001])
def test_percentile_bit_equal_numpy(torch_cuda, count):
    """Order statistics by selection (register-resident keys up to 32768
    values per series, the streaming variant beyond) are bit-equal to
    np.percentile, for every keys-per-thread instantiation."""
    rs = np.random.RandomState(count)
    D = rs.randn(count, 7, 3)
    eng = make_engine('behavioral', 4, 8, 1, [4])
    lo, hi = eng.percentile(D, 2.5, 97.5)
    want = np.percentile(D, [2.5, 97.5], axis=0)
    assert np.array_equal(lo.cpu().numpy(), want[0])
    assert np.array_equal(hi.cpu().numpy(), want[1])


def test_percentile_ties_signs_nan_and_other_quantiles(torch_cuda):
    rs = np.random.RandomState(5)
    eng = make_engine('behavioral', 4, 8, 1, [4])
    D = rs.randn(999, 6)
    D[:, 1] = np.round(D[:, 1])                # heavy ties, both signs, +-0
    D[:, 2] = -np.abs(D[:, 2])                 # all negative
    D[:, 3] = 7.25                             # constant series
    D[::3, 4] = 0.0
    D[1::3, 4] = -0.0
    for qlo, qhi in ((2.5, 97.5), (0.0, 100.0), (50.0, 99.9), (0.5, 5.0)):
        lo, hi = eng.percentile(D, qlo, qhi)
        want = np.percentile(D, [qlo, qhi], axis=0)
        assert np.array_equal(lo.cpu().numpy(), want[0])
        assert np.array_equal(hi.cpu().numpy(), want[1])
    D[17, 5] = np.nan                          # np.percentile propagates NaN
    lo, hi = eng.percentile(D, 2.5, 97.5)
    assert np.isnan(lo.cpu().numpy()[5]) and np.isnan(hi.cpu().numpy()[5])
    assert np.array_equal(lo.cpu().numpy()[:5],
                          np.percentile(D[:, :5], 2.5, axis=0))


def test_percentile_series_major_and_transpose(torch_cuda):
    """The series-major entry (what a rank holds after the series exchange of
    a multi-GPU run), padded rows included, and the transpose that feeds it."""
    import torch
    rs = np.random.RandomState(6)
    eng = make_engine('behavioral', 4, 8, 1, [4])
    D = rs.randn(1234, 37)
    Dt = eng.transpose(D)
    assert np.array_equal(Dt.cpu().numpy(), D.T)
    want = np.percentile(D, [2.5, 97.5], axis=0)
    lo, hi = eng.percentile_series(Dt, 2.5, 97.5)
    assert np.array_equal(lo.cpu().numpy(), want[0])
    assert np.array_equal(hi.cpu().numpy(), want[1])
    padded = torch.full((37, 1300), 1e300, dtype=torch.float64, device='cuda')
    padded[:, :1234] = Dt
    lo, hi = eng.percentile_series(padded[:, :1234], 2.5, 97.5)
    assert np.array_equal(lo.cpu().numpy(), want[0])
    assert np.array_equal(hi.cpu().numpy(), want[1])


def test_device_gaussian_tables_replay_numpy_legacy_stream(torch_cuda):
    """Table i = RandomState(i).normal(size=(T, 11)) (what the reference's
    compute.svd(..., seed=i) makes sklearn draw): the MT19937 uniforms are
    replayed bit for bit, the polar-method deviates up to the last bit of the
    device's log / sqrt."""
    from pypyls_b200.types.regression import gaussian_tables
    eng = make_engine('regression', 30, 40, 20, [30], n_components=3)
    for first, count in ((0, 70), (4999, 33), (2 ** 31 + 5, 4)):
        got = eng.gen_gaussian_tables(first, count).cpu().numpy()
        want = gaussian_tables(range(first, first + count), 20)
        np.testing.assert_allclose(got, want, rtol=4e-16, atol=0)
        assert np.mean(got == want) > 0.9


def test_pvals_and_boot_ratio(torch_cuda):
    rs = np.random.RandomState(11)
    eng = make_engine('behavioral', 4, 8, 1, [4])
    dperm, dorig = rs.rand(777, 9), rs.rand(9)
    dperm[5, 2] = dorig[2]                     # ties do not count (strict >)
    p = eng.perm_pvals(dperm, dorig).cpu().numpy()
    assert np.array_equal(p, po.perm_sig(np.diag(dorig), dperm.T))
    bs, us, uq = rs.randn(50, 4), rs.randn(50, 4) * 30, rs.rand(50, 4) * 900
    import torch
    for add, n in ((True, 101), (False, 100)):
        bsr, se = eng.boot_ratio(bs, torch.from_numpy(us).cuda(),
                                 torch.from_numpy(uq).cuda(), 100, add)
        a, b = (us + bs, uq + bs ** 2) if add else (us, uq)
        want_bsr, want_se = po.boot_rel(bs, a, b, n)
        np.testing.assert_allclose(se.cpu().numpy(), want_se, rtol=1e-14)
        np.testing.assert_allclose(bsr.cpu().numpy(), want_bsr, rtol=1e-14)


# ---------------------------------------------------------------------------
LAYOUTS = [([20], 1), ([10, 10], 2), ([7, 9, 5], 3), ([12], 2), ([4, 3], 1)]


@pytest.mark.parametrize('groups,n_cond', LAYOUTS)
@pytest.mark.parametrize('mode', ['behavioral', 'behavioral_cov'])
def test_crosscov_behavioral(torch_cuda, groups, n_cond, mode):
    """Cross-covariance of permuted / bootstrapped data == the reference's
    gather-then-xcorr (pyls/base.py:569,599; behavioral.py:27-52)."""
    rs = np.random.RandomState(sum(groups) + n_cond)
    S, B, T = sum(groups) * n_cond, 150, 3
    X, Y = rs.rand(S, B) + 5.0, rs.rand(S, T)
    spec = po._Spec('behavioral', groups, n_cond,
                    covariance=(mode == 'behavioral_cov'))
    eng = make_engine(mode, S, B, T, groups, n_cond).set_data(X, Y)
    R0 = eng.crosscov().cpu().numpy()[0]
    np.testing.assert_allclose(R0, po.gen_covcorr(spec, X, Y), rtol=1e-10,
                               atol=1e-13)
    perms = po.gen_permsamp(groups, n_cond, 6, seed=1)
    Rp = eng.crosscov(perms).cpu().numpy()
    boots = po.gen_bootsamp(groups, n_cond, 6, seed=2)
    Rb = eng.crosscov(boots, bootstrap=True).cpu().numpy()
    for i in range(6):
        np.testing.assert_allclose(
            Rp[i], po.gen_covcorr(spec, X, Y[perms[:, i]]), rtol=1e-10,
            atol=1e-13)
        np.testing.assert_allclose(
            Rb[i], po.gen_covcorr(spec, X[boots[:, i]], Y[boots[:, i]]),
            rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('groups,n_cond', [([10, 10], 2), ([7, 9, 5], 3),
                                           ([12], 2), ([4, 3], 1)])
@pytest.mark.parametrize('mc', [0, 1, 2])
def test_crosscov_meancentered(torch_cuda, groups, n_cond, mc):
    if (n_cond == 1 and mc == 0) or (len(groups) == 1 and mc == 1):
        pytest.skip('combination rewritten by the front-end')
    rs = np.random.RandomState(sum(groups) + n_cond + mc)
    S, B = sum(groups) * n_cond, 140
    X = rs.rand(S, B) + 3.0
    spec = po._Spec('meancentered', groups, n_cond, mean_centering=mc)
    eng = make_engine('meancentered', S, B, 1, groups, n_cond,
                      mean_centering=mc).set_data(X)
    np.testing.assert_allclose(eng.crosscov().cpu().numpy()[0],
                               po.gen_covcorr(spec, X, spec.dummy),
                               rtol=1e-10, atol=1e-13)
    perms = po.gen_permsamp(groups, n_cond, 5, seed=3)
    boots = po.gen_bootsamp(groups, n_cond, 5, seed=4)
    Rp = eng.crosscov(perms).cpu().numpy()
    Rb = eng.crosscov(boots, bootstrap=True).cpu().numpy()
    for i in range(5):
        np.testing.assert_allclose(
            Rp[i], po.gen_covcorr(spec, X[perms[:, i]], spec.dummy),
            rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(
            Rb[i], po.gen_covcorr(spec, X[boots[:, i]], spec.dummy),
            rtol=1e-10, atol=1e-13)


# ---------------------------------------------------------------------------
@pytest.mark.parametrize('groups,n_cond', LAYOUTS)
def test_device_index_tables_obey_reference_invariants(torch_cuda, groups,
                                                       n_cond):
    """Acceptance test of the on-device generator: the invariants the
    reference pins in pyls/tests/test_base.py:14-115."""
    S, n_subj, n = sum(groups) * n_cond, sum(groups), 200
    if groups == [4, 3]:
        n = 12                       # tiny layout: few distinct resamples exist
    eng = make_engine('meancentered' if len(groups) * n_cond > 1
                      else 'behavioral', S, 64, 1, groups, n_cond)
    perm, _ = eng.gen_perm_indices(1234, n)
    boot, n_ex = eng.gen_boot_indices(1234, n)
    perm, boot = perm.cpu().numpy().T, boot.cpu().numpy().T
    assert perm.shape == boot.shape == (S, n)
    # every permutation column is a permutation of the rows; columns unique
    assert np.all(np.sort(perm, axis=0) == np.arange(S)[:, None])
    assert len({c.tobytes() for c in perm.T}) == n
    # same seed -> same table; other seed -> other table; prefix property
    again, _ = eng.gen_perm_indices(1234, n)
    assert np.array_equal(again.cpu().numpy().T, perm)
    other, _ = eng.gen_perm_indices(4321, n)
    assert not np.array_equal(other.cpu().numpy().T, perm)
    tail, _ = eng.gen_perm_indices(1234, n - 5, first=5)
    assert np.array_equal(tail.cpu().numpy().T, perm[:, 5:])
    tailb, _ = eng.gen_boot_indices(1234, n - 5, first=5)
    assert np.array_equal(tailb.cpu().numpy().T, boot[:, 5:])

    bounds = np.concatenate([[0], np.cumsum(groups)])
    subj_of_row = np.concatenate([
        np.tile(np.arange(a, b), n_cond)
        for a, b in zip(bounds[:-1], bounds[1:])])
    cond_of_row = np.concatenate([
        np.repeat(np.arange(n_cond), b - a)
        for a, b in zip(bounds[:-1], bounds[1:])])
    grp_of_subj = np.repeat(np.arange(len(groups)), groups)
    for table, is_boot in ((perm, False), (boot, True)):
        for col in table.T:
            src_subj = subj_of_row[col]
            # conditions of one destination subject come from ONE source
            # subject (test_base.py:38-46, 110-115)
            for a, b in zip(bounds[:-1], bounds[1:]):
                rows = np.where((subj_of_row >= a) & (subj_of_row < b))[0]
                per_cond = src_subj[rows].reshape(n_cond, b - a)
                assert np.all(per_cond == per_cond[0])
                if is_boot:
                    # resampled inside the group, sorted, conditions in place
                    assert np.all(grp_of_subj[per_cond[0]] ==
                                  grp_of_subj[a])
                    assert np.all(np.diff(per_cond[0]) >= 0)
                    assert np.all(cond_of_row[col[rows]] == cond_of_row[rows])
                    assert np.unique(per_cond[0]).size >= \
                        int(np.ceil(min(groups) / 2))
                else:
                    # every condition exactly once per destination subject
                    conds = cond_of_row[col[rows]].reshape(n_cond, b - a)
                    assert np.all(np.sort(conds, axis=0) ==
                                  np.arange(n_cond)[:, None])
            if not is_boot and len(groups) > 1:
                # subjects always mix across groups (test_base.py:50-54)
                for a, b in zip(bounds[:-1], bounds[1:]):
                    rows = np.where((subj_of_row >= a) & (subj_of_row < b))[0]
                    assert not np.all(np.isin(src_subj[rows],
                                              np.arange(a, b)))
    if n_ex == 0:
        for a, b in zip(bounds[:-1], bounds[1:]):
            assert len({c[a:b].tobytes() for c in boot.T}) == n


def test_device_permutations_are_uniform(torch_cuda):
    """Every subject should land on every slot about equally often."""
    eng = make_engine('behavioral', 8, 16, 1, [8])
    perm, _ = eng.gen_perm_indices(7, 4000)
    perm = perm.cpu().numpy()
    counts = np.stack([np.bincount(perm[:, s], minlength=8)
                       for s in range(8)])
    expected = 4000 / 8
    chi2 = ((counts - expected) ** 2 / expected).sum()
    assert chi2 < 120          # 49 dof; P(chi2 > 120) ~ 1e-7
