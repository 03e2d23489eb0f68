# -*- coding: utf-8 -*-
"""
CPU oracle for the PLS resampling hot path -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the per-resample algorithm of
netneurolab/pypyls (reference checked out at /root/reference, commit e0ff056).
It exists so that the CUDA engine in ``pypyls_b200`` can be checked against the
reference's arithmetic on a box where the reference itself is not present.

Rules (enforced by
tests/test_host_logic.py::test_product_never_imports_the_oracle):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this module;
  * the product package ``pypyls_b200`` never imports it and has no CPU
    fallback.

Parity status: PINNED.  ``tests/golden/make_golden.py`` (run in the build
container, where /root/reference is importable with three shims) stores the
reference's own outputs for several small analyses plus the in-tree Matlab
fixtures; ``tests/test_oracle_golden.py`` checks every function here against
them.  The exception is SIMPLS resampling (``pls_regression``): the reference's
tests only pin shapes for it, so the oracle is pinned against the shimmed
reference run only ("parity unpinned by the reference's tests").

Third-party arithmetic the reference calls and this oracle calls identically
(not vendored by the reference; versions in this image are the de-facto pin):
``sklearn.utils.extmath.randomized_svd`` (scikit-learn 1.9.0),
``scipy.stats.zscore`` (scipy 1.18.1), ``numpy.percentile`` (numpy 2.3.5).

Every function cites the reference file:line it follows.
"""

import warnings

import numpy as np
from scipy.stats import zscore
from sklearn.utils.extmath import randomized_svd
from sklearn.utils.validation import check_random_state


# --------------------------------------------------------------------------
# layout helpers
# --------------------------------------------------------------------------

def cell_labels(groups, n_cond=1):
    """Cell label (1-based) of every row; rows are ordered group -> condition
    -> subject.  Follows pyls/utils.py:178-197 (dummy_label)."""
    groups = [int(g) for g in groups]
    n_cells = len(groups) * n_cond
    return np.repeat(np.arange(1, n_cells + 1), np.repeat(groups, n_cond))


def dummy_code(groups, n_cond=1):
    """(S, J) 0/1 cell-membership matrix.  Follows pyls/utils.py:155-175."""
    lab = cell_labels(groups, n_cond)
    return (lab[:, None] == np.unique(lab)[None, :]).astype(int)


def permute_cols(x, rs):
    """Shuffle the rows of every column independently by arg-sorting uniform
    draws.  Follows pyls/utils.py:200-224."""
    x = np.asarray(x)
    order = rs.random_sample(x.shape).argsort(axis=0)
    return x[order, np.arange(x.shape[1])[None, :]]


# --------------------------------------------------------------------------
# index generators (integer work -> bit-exact with the reference for a seed)
# --------------------------------------------------------------------------

def _cond_index_blocks(groups, n_cond):
    """Per group, the (n_cond, n_subj) array of row ids (row c = condition c).
    Restates the np.where/np.split construction of pyls/base.py:42-45."""
    blocks, start = [], 0
    for g in groups:
        blocks.append(start + np.arange(n_cond * g).reshape(n_cond, g))
        start += n_cond * g
    return blocks


def _stack_by_group(table, perm, groups):
    """Take subject columns ``perm`` of the (n_cond, n_subj_total) ``table``,
    cut them into the groups' sizes and lay every group out condition-major.
    Restates pyls/base.py:64-65 and :142-143."""
    picked = table[:, perm].T
    out, start = [], 0
    for g in groups:
        out.append(picked[start:start + g].flatten('F'))
        start += g
    return np.hstack(out)


def gen_permsamp(groups, n_cond, n_perm, seed=None):
    """Permutation index table (S, n_perm).  Follows pyls/base.py:10-79:
    per-subject condition shuffle, cross-group subject permutation, rejection
    when a group keeps its own subject set or the column repeats an earlier
    one, give up (warn once) after 500 tries."""
    groups = [int(g) for g in groups]
    n_rows = sum(groups) * n_cond
    n_subj = sum(groups)
    rs = check_random_state(seed)
    blocks = _cond_index_blocks(groups, n_cond)
    bounds = np.concatenate([[0], np.cumsum(groups)])
    out = np.zeros((n_rows, n_perm), dtype=int)
    subj = np.arange(n_subj)
    warned = False
    for i in range(n_perm):
        tries, bad = 0, True
        while bad and tries < 500:
            tries, bad = tries + 1, False
            table = np.hstack([permute_cols(b, rs) for b in blocks])
            perm = rs.permutation(subj)
            if len(groups) > 1:
                for a, b in zip(bounds[:-1], bounds[1:]):
                    if np.array_equal(np.sort(perm[a:b]), subj[a:b]):
                        bad = True
            col = _stack_by_group(table, perm, groups)
            if i and (col[:, None] == out[:, :i]).all(axis=0).any():
                bad = True
        if tries == 500 and not warned:
            warnings.warn('WARNING: Duplicate permutations used.')
            warned = True
        out[:, i] = col
    return out


def gen_bootsamp(groups, n_cond, n_boot, seed=None):
    """Bootstrap index table (S, n_boot).  Follows pyls/base.py:82-159: sorted
    within-group sampling with replacement, at least ceil(min_cell/2) distinct
    subjects per group, conditions follow their subject, per-group duplicate
    rejection (the reference compares rows indexed by SUBJECT id, :145-149 --
    reproduced as is), give up after 500 tries."""
    groups = [int(g) for g in groups]
    n_rows = sum(groups) * n_cond
    n_subj = sum(groups)
    rs = check_random_state(seed)
    min_subj = int(np.ceil(min(groups) * 0.5))
    table = np.hstack(_cond_index_blocks(groups, n_cond))
    bounds = np.concatenate([[0], np.cumsum(groups)])
    out = np.zeros((n_rows, n_boot), dtype=int)
    warned = False
    for i in range(n_boot):
        tries, bad = 0, True
        while bad and tries < 500:
            tries, bad = tries + 1, False
            boot = np.zeros(n_subj, dtype=int)
            for a, b in zip(bounds[:-1], bounds[1:]):
                pool = np.arange(a, b)
                while True:
                    boot[a:b] = np.sort(rs.choice(pool, size=b - a,
                                                  replace=True))
                    if np.unique(boot[a:b]).size >= min_subj:
                        break
            col = _stack_by_group(table, boot, groups)
            for a, b in zip(bounds[:-1], bounds[1:]):
                if i and (col[a:b, None] == out[a:b, :i]).all(axis=0).any():
                    bad = True
        if tries == 500 and not warned:
            warnings.warn('WARNING: Duplicate bootstraps used.')
            warned = True
        out[:, i] = col
    return out


def gen_splits(groups, n_cond, n_split, seed=None, test_size=0.5):
    """Train / split-half masks (S, n_split) bool.  Follows pyls/base.py:162-229:
    per group a coin flip between ceil and floor of n_g * (1 - test_size)
    subjects drawn without replacement, conditions follow their subject,
    duplicate columns rejected, give up (warn once) after 500 tries."""
    groups = [int(g) for g in groups]
    n_subj = sum(groups)
    rs = check_random_state(seed)
    bounds = np.concatenate([[0], np.cumsum(groups)])
    out = np.zeros((n_subj * n_cond, n_split), dtype=bool)
    warned = False
    for i in range(n_split):
        tries, bad = 0, True
        while bad and tries < 500:
            tries, bad = tries + 1, False
            split = np.zeros(n_subj, dtype=bool)
            for a, b in zip(bounds[:-1], bounds[1:]):
                take = rs.choice([np.ceil, np.floor])
                num = int(take((b - a) * (1 - test_size)))
                split[rs.choice(np.arange(a, b), size=num, replace=False)] = True
            # rows: groups stacked, conditions within group, subjects within
            half = np.hstack([np.tile(split[a:b], n_cond)
                              for a, b in zip(bounds[:-1], bounds[1:])])
            if i and (half[:, None] == out[:, :i]).all(axis=0).any():
                bad = True
        if tries == 500 and not warned:
            warnings.warn('WARNING: Duplicate split halves used.')
            warned = True
        out[:, i] = half
    return out


# --------------------------------------------------------------------------
# numeric primitives (pyls/compute.py)
# --------------------------------------------------------------------------

def svd(crosscov, n_components=None, seed=None):
    """Randomized SVD on the tall orientation; returns (U (B,L), diag(d), V).
    Follows pyls/compute.py:10-52."""
    rs = check_random_state(seed)
    crosscov = np.asanyarray(crosscov)
    if n_components is None:
        n_components = min(crosscov.shape)
    if crosscov.shape[0] <= crosscov.shape[1]:
        U, d, Vt = randomized_svd(crosscov.T, n_components=n_components,
                                  random_state=rs, transpose=False)
        V = Vt.T
    else:
        V, d, Ut = randomized_svd(crosscov, n_components=n_components,
                                  random_state=rs, transpose=False)
        U = Ut.T
    return U, np.diag(d), V


def xcorr(X, Y, covariance=False):
    """(T, B) cross-correlation (z-scored, ddof=1) or cross-covariance of the
    rows given.  Follows pyls/compute.py:55-94 (norm=False branch)."""
    if not covariance:
        Xn = (X - X.mean(axis=0)) / X.std(axis=0, ddof=1)
        Yn = (Y - Y.mean(axis=0)) / Y.std(axis=0, ddof=1)
    else:
        Xn = X - X.mean(0, keepdims=True)
        Yn = Y - Y.mean(0, keepdims=True)
    return (Yn.T @ Xn) / (len(Xn) - 1)


def normalize(X, axis=0):
    """Unit-norm columns, zero-safe.  Follows pyls/compute.py:97-126."""
    out = np.array(X, dtype=float)
    nrm = np.linalg.norm(out, axis=axis, keepdims=True)
    zero = nrm == 0
    nrm[zero] = 1
    out = out / nrm
    out[np.broadcast_to(zero, out.shape)] = 0
    return out


def rescale_test(X_train, X_test, Y_train, U, V):
    """Out-of-sample prediction of Y.  pyls/compute.py:129-151 (scipy's zmap
    written out: test columns standardised with the training mean / ddof=1
    standard deviation)."""
    mu = X_train.mean(axis=0, keepdims=True)
    sd = X_train.std(axis=0, ddof=1, keepdims=True)
    return ((X_test - mu) / sd) @ U @ V.T + Y_train.mean(axis=0, keepdims=True)


def r2_score_raw(y_true, y_pred):
    """sklearn.metrics.r2_score(..., multioutput='raw_values') for columns
    with non-constant truth (pyls/types/behavioral.py:168)."""
    num = ((y_true - y_pred) ** 2).sum(axis=0)
    den = ((y_true - y_true.mean(axis=0)) ** 2).sum(axis=0)
    return 1 - num / den


def perm_sig(orig, perm):
    """(#[perm > orig] + 1) / (P + 1), strict '>'.  pyls/compute.py:154-181."""
    count = np.sum(perm > np.diag(orig)[:, None], axis=1) + 1
    return count / (perm.shape[-1] + 1)


def boot_ci(boot, ci=95):
    """Percentile CI along the last axis.  pyls/compute.py:184-209."""
    low = (100 - ci) / 2
    lower, upper = np.percentile(boot, [low, 100 - low], axis=-1)
    return lower, upper


def boot_rel(orig, u_sum, u_square, n_boot):
    """Bootstrap ratio and standard error.  pyls/compute.py:212-237."""
    u_sum2 = (u_sum ** 2) / n_boot
    u_se = np.sqrt(np.abs(u_square - u_sum2) / (n_boot - 1))
    with np.errstate(divide='ignore', invalid='ignore'):
        bsr = orig / u_se
    return bsr, u_se


def procrustes(original, permuted, singular):
    """Rotate ``permuted @ singular`` onto ``original``: polar factor of
    original.T @ permuted.  Follows pyls/compute.py:240-264 (the reference
    calls randomized_svd with the global RNG; full-rank so any seed gives the
    same factor up to rounding -- a fixed seed is used here)."""
    temp = original.T @ permuted
    N, _, P = randomized_svd(temp, n_components=min(temp.shape),
                             random_state=0)
    return permuted @ singular @ (P.T @ N.T)


def get_group_mean(X, dummy, n_cond=1, mean_centering=0):
    """(J, B) mean to remove from every cell.  pyls/compute.py:267-317."""
    J = dummy.shape[-1]
    if mean_centering == 0:
        sizes = dummy[:, 0:J:n_cond].sum(axis=0).astype(int) * n_cond
        member = dummy_code(sizes)
    elif mean_centering == 1:
        member = dummy.copy()
    elif mean_centering == 2:
        member = np.ones((len(X), 1))
    else:
        raise ValueError("Mean centering type must be in [0, 1, 2].")
    means = np.vstack([X[m].mean(axis=0)[None]
                       for m in member.T.astype(bool)])
    if mean_centering == 0:
        means = np.repeat(means, n_cond, axis=0)
    elif mean_centering == 1:
        means = means.reshape(-1, n_cond, X.shape[-1]).mean(axis=0)
        means = np.tile(means.T, int(J / n_cond)).T
    else:
        means = np.repeat(means, J, axis=0)
    return means


def get_mean_center(X, dummy, n_cond=1, mean_centering=0, means=True):
    """Cell means minus the centering mean ((J, B), means=True) or de-meaned
    rows ((S, B)).  Follows pyls/compute.py:320-357."""
    mc = get_group_mean(X, dummy, n_cond=n_cond, mean_centering=mean_centering)
    cells = dummy.T.astype(bool)
    if means:
        return np.vstack([X[c].mean(axis=0) - mc[n]
                          for n, c in enumerate(cells)])
    return np.vstack([X[c] - mc[n][None] for n, c in enumerate(cells)])


def efficient_corr(x, y):
    """Column-wise Pearson r, clipped.  pyls/compute.py:360-391."""
    x, y = np.vstack(x), np.vstack(y)
    corr = np.sum(zscore(x, ddof=1) * zscore(y, ddof=1), axis=0) / (len(x) - 1)
    return np.clip(corr, -1, 1)


def varexp(singular):
    """Squared singular values normalised to 1.  pyls/compute.py:394-414."""
    sq = np.diag(singular) ** 2
    return np.diag(sq / np.sum(sq))


# --------------------------------------------------------------------------
# type-specific cross-covariance / distribution builders
# --------------------------------------------------------------------------

class _Spec:
    """What kind of analysis is running (replaces the reference's subclasses
    of BasePLS)."""

    def __init__(self, kind, groups, n_cond, covariance=False,
                 mean_centering=0, rotate=True, n_components=None):
        self.kind = kind                    # 'behavioral' | 'meancentered' | 'regression'
        self.groups = [int(g) for g in groups]
        self.n_cond = int(n_cond)
        self.covariance = covariance
        self.mean_centering = mean_centering
        self.rotate = rotate
        self.n_components = n_components
        self.dummy = dummy_code(self.groups, self.n_cond)


def gen_covcorr(spec, X, Y, dummy=None):
    """Matrix that is decomposed.  behavioral: row-stack of per-cell xcorr
    (pyls/types/behavioral.py:27-52); mean-centered: cell means minus
    centering mean (pyls/types/meancentered.py:50-73)."""
    dummy = spec.dummy if dummy is None else dummy
    if spec.kind == 'behavioral':
        return np.vstack([xcorr(X[c], Y[c], covariance=spec.covariance)
                          for c in dummy.T.astype(bool)])
    return get_mean_center(X, dummy, spec.n_cond, spec.mean_centering,
                           means=True)


def gen_distrib(spec, X, Y, original, dummy=None):
    """Bootstrap distribution entry.  behavioral: per-cell xcorr of the scores
    X @ normalize(U_orig) with Y (pyls/types/behavioral.py:54-80);
    mean-centered: cell means of the de-meaned rows projected on
    normalize(U_orig) (pyls/types/meancentered.py:75-102)."""
    dummy = spec.dummy if dummy is None else dummy
    if spec.kind == 'behavioral':
        return gen_covcorr(spec, X @ normalize(original), Y, dummy)
    usc = get_mean_center(X, dummy, spec.n_cond, spec.mean_centering,
                          means=False)
    usc = usc @ normalize(original)
    return np.vstack([usc[c].mean(axis=0) for c in dummy.T.astype(bool)])


def decompose(spec, X, Y, seed=None):
    """gen_covcorr followed by the SVD.  pyls/base.py:401-437; regression:
    pyls/types/regression.py:248-277."""
    if spec.kind == 'regression':
        mask = get_mask(X, Y)
        out = simpls(X[mask], Y[mask], spec.n_components, seed=seed)
        return out['x_weights'], np.diag(out['pctvar'][1]), None
    return svd(gen_covcorr(spec, X, Y), seed=seed)


# --------------------------------------------------------------------------
# per-resample bodies (pyls/base.py)
# --------------------------------------------------------------------------

def split_half(spec, X, Y, ud, vd, n_split, seed=None, splits=None):
    """Split-half reliability of the singular vectors -> (ucorr (L,), vcorr
    (L,)).  Follows pyls/base.py:714-770: ``n_split`` half/half masks from
    gen_splits(seed, test_size=0.5); for every mask the cross-covariance of
    either half is projected on the scaled singular vectors ``vd = V d^-1``
    (-> B x L) and ``ud = U d^-1`` (-> K x L) and matching columns of the two
    halves are correlated; the correlations are averaged over the masks."""
    if splits is None:
        splits = gen_splits(spec.groups, spec.n_cond, n_split, seed=seed,
                            test_size=0.5)
    dummy = spec.dummy
    ucorr = np.zeros((ud.shape[-1], n_split))
    vcorr = np.zeros((vd.shape[-1], n_split))
    for i in range(n_split):
        spl = splits[:, i].astype(bool)
        D1 = gen_covcorr(spec, X[spl], Y[spl], dummy[spl])
        D2 = gen_covcorr(spec, X[~spl], Y[~spl], dummy[~spl])
        ucorr[:, i] = efficient_corr(D1.T @ vd, D2.T @ vd)
        vcorr[:, i] = efficient_corr(D1 @ ud, D2 @ ud)
    return ucorr.mean(axis=-1), vcorr.mean(axis=-1)


def single_perm(spec, X, Y, perminds, original_v, seed=None, use_permind=True,
                n_split=None):
    """One permutation -> (L,) permuted singular values (and, with
    ``n_split``, the split-half correlations of the permuted data).  Follows
    pyls/base.py:654-712; the permuted operand is Y for
    behavioral (base.py:599) and X for mean-centered
    (pyls/types/meancentered.py:125).  use_permind=False: ``perminds`` is a
    pre-permuted Y matrix used as it is (base.py:689-692)."""
    if spec.kind == 'regression':
        return _regression_single_perm(spec, X, Y, perminds, seed)
    if not use_permind:
        Xp, Yp = X, perminds
    elif spec.kind == 'meancentered':
        Xp, Yp = X[perminds], Y
    else:
        Xp, Yp = X, Y[perminds]
    U, d, V = decompose(spec, Xp, Yp, seed=seed)
    if spec.rotate:
        rot = procrustes(original_v, V, d)
        ssd = np.sqrt(np.sum(rot ** 2, axis=0))
    else:
        ssd = np.diag(d)
    if n_split is None:
        return ssd
    # pyls/base.py:704-708: the permutation's own (un-rotated) decomposition,
    # masks seeded by the permutation number
    di = np.linalg.inv(d)
    ucorr, vcorr = split_half(spec, Xp, Yp, U @ di, V @ di, n_split, seed=seed)
    return ssd, ucorr, vcorr


def single_boot(spec, X, Y, inds, original_u, seed=None):
    """One bootstrap -> (distrib (K, L), U_boot (B, L)).  Follows
    pyls/base.py:530-576."""
    if spec.kind == 'regression':
        return _regression_single_boot(spec, X, Y, inds, original_u, seed)
    Xb, Yb = X[inds], Y[inds]
    U, d = decompose(spec, Xb, Yb, seed=seed)[:-1]
    U_boot = procrustes(original_u, U, d)
    distrib = gen_distrib(spec, Xb, Yb, original_u)
    return distrib, U_boot


def single_boot_nullsafe(spec, X, Y, inds, original_u, d_orig):
    """One bootstrap with numerically null latent variables LEFT OUT of the
    Procrustes rotation -- the documented deviation of the CUDA engine from
    pyls/base.py:530-576 for rank-deficient cross-covariances (mean-centred
    PLS: rank J-1).  The reference rotates with the arbitrary unit vectors its
    randomized SVD returns for the null directions (rounding noise, different
    for every BLAS), which perturbs the non-null latent variables by
    O(sqrt(K/B)); here the rotation is the Procrustes solution restricted to
    the non-null subspaces and the null columns of U_boot are 0.  Uses an
    exact LAPACK SVD.  For full-rank problems this equals single_boot."""
    Xb, Yb = X[inds], Y[inds]
    R = gen_covcorr(spec, Xb, Yb)
    U, d, _ = np.linalg.svd(R.T, full_matrices=False)
    keep_b = d > 1e-7 * d.max()
    keep_o = d_orig > 1e-10 * d_orig.max()
    temp = original_u[:, keep_o].T @ U[:, keep_b]
    N, _, P = np.linalg.svd(temp, full_matrices=False)
    U_boot = np.zeros_like(original_u)
    U_boot[:, keep_o] = (U[:, keep_b] * d[keep_b]) @ (P.T @ N.T)
    return gen_distrib(spec, Xb, Yb, original_u), U_boot


def run_boots_nullsafe(spec, X, Y, bootsamp, original_u, d_orig):
    """run_boots over single_boot_nullsafe."""
    u_sum = np.zeros_like(original_u)
    u_square = np.zeros_like(original_u)
    distrib = []
    for i in range(bootsamp.shape[-1]):
        d, u = single_boot_nullsafe(spec, X, Y, bootsamp[:, i], original_u,
                                    d_orig)
        u_sum += u
        u_square += u ** 2
        distrib.append(d)
    return np.stack(distrib, axis=-1), u_sum, u_square


def run_perms(spec, X, Y, permsamp, original_v, first=0, count=None,
              use_permind=True):
    """d_perm (L, count) over [first, first+count) of the last axis of
    ``permsamp`` -- (S, P) index vectors or, with use_permind=False, (S, T, P)
    pre-permuted Y matrices; resample i uses seed=i exactly like
    pyls/base.py:644-650."""
    count = permsamp.shape[-1] - first if count is None else count
    cols = [single_perm(spec, X, Y, permsamp[..., i], original_v, seed=i,
                        use_permind=use_permind)
            for i in range(first, first + count)]
    return np.stack(cols, axis=-1)


def run_perms_split(spec, X, Y, permsamp, original_v, n_split, first=0,
                    count=None, use_permind=True):
    """run_perms with split-half resampling inside every permutation ->
    (d_perm (L, count), ucorrs (L, count), vcorrs (L, count));
    pyls/base.py:644-652 with n_split set."""
    count = permsamp.shape[-1] - first if count is None else count
    out = [single_perm(spec, X, Y, permsamp[..., i], original_v, seed=i,
                       use_permind=use_permind, n_split=n_split)
           for i in range(first, first + count)]
    return tuple(np.stack([o[k] for o in out], axis=-1) for k in range(3))


def _split_results(spec, X, Y, res, U, d, V, n_split, ucorrs, vcorrs, rs, ci):
    """Split-half block of BasePLS.run_pls (pyls/base.py:373-397): the
    original data's correlations (masks drawn from the analysis' RandomState
    after the permutation table), their p-values against the permuted
    correlations and percentile limits of the latter."""
    di = np.linalg.inv(d)
    ou, ov = split_half(spec, X, Y, U @ di, V @ di, n_split, seed=rs)
    res['ucorr'], res['vcorr'] = ou, ov
    res['ucorr_pvals'] = perm_sig(np.diag(ou), ucorrs)
    res['vcorr_pvals'] = perm_sig(np.diag(ov), vcorrs)
    res['ucorr_lolim'], res['ucorr_uplim'] = boot_ci(ucorrs, ci=ci)
    res['vcorr_lolim'], res['vcorr_uplim'] = boot_ci(vcorrs, ci=ci)
    res['perm_ucorr'], res['perm_vcorr'] = ucorrs, vcorrs


def single_crossval(spec, X, Y, inds, seed=None):
    """One train / test split -> (r (T,), r2 (T,)).  Follows
    BehavioralPLS._single_crossval, pyls/types/behavioral.py:125-170."""
    dummy = spec.dummy
    Xtr, Ytr, dtr = X[inds], Y[inds], dummy[inds]
    Xte, Yte, dte = X[~inds], Y[~inds], dummy[~inds]
    U, d, V = svd(gen_covcorr(spec, Xtr, Ytr, dummy=dtr), seed=seed)
    pred = []
    for n, V_spl in enumerate(np.split(V, dummy.shape[-1])):
        tr, te = dtr[:, n].astype(bool), dte[:, n].astype(bool)
        pred.append(rescale_test(Xtr[tr], Xte[te], Ytr[tr], U, V_spl))
    pred = np.vstack(pred)
    return efficient_corr(Yte, pred), r2_score_raw(Yte, pred)


def crossval(spec, X, Y, n_split, test_size=0.25, seed=None, splits=None):
    """(r_scores (T, C), r2_scores (T, C)) over C train / test splits.
    BehavioralPLS.crossval, pyls/types/behavioral.py:82-123."""
    if splits is None:
        splits = gen_splits(spec.groups, spec.n_cond, n_split, seed=seed,
                            test_size=test_size)
    out = [single_crossval(spec, X, Y, splits[:, i], seed=i)
           for i in range(splits.shape[1])]
    r, r2 = [np.stack(o, axis=-1) for o in zip(*out)]
    return r, r2


def run_boots(spec, X, Y, bootsamp, original_u, first=0, count=None):
    """(distrib (K, L, count), u_sum, u_square) as pyls/base.py:439-528."""
    count = bootsamp.shape[-1] - first if count is None else count
    u_sum = np.zeros_like(original_u)
    u_square = np.zeros_like(original_u)
    distrib = []
    for i in range(first, first + count):
        d, u = single_boot(spec, X, Y, bootsamp[:, i], original_u, seed=i)
        u_sum += u
        u_square += u ** 2
        distrib.append(d)
    return np.stack(distrib, axis=-1), u_sum, u_square


# --------------------------------------------------------------------------
# SIMPLS (pyls/types/regression.py)
# --------------------------------------------------------------------------

def get_mask(X, Y):
    """Rows where neither X nor Y is all-NaN.  regression.py:48-53."""
    return ~(np.all(np.isnan(X), axis=1) | np.all(np.isnan(Y), axis=1))


def simpls(X, Y, n_components=None, seed=1234):
    """SIMPLS with the reference's randomized top-1 SVD per component, double
    modified Gram-Schmidt and deflation.  Follows
    pyls/types/regression.py:56-186; only the outputs the resampling path
    consumes (x_weights, pctvar, loadings, scores) are produced -- the
    reference's mse / t2 / residual reconstructions (:159-172) are unused by
    the permutation and bootstrap loops and are omitted."""
    X, Y = np.asanyarray(X), np.asanyarray(Y)
    if n_components is None:
        n_components = min(len(X) - 1, X.shape[1])
    X0 = X - X.mean(axis=0, keepdims=True)
    Y0 = Y - Y.mean(axis=0, keepdims=True)
    Cov = X0.T @ Y0
    B, T, S = X.shape[1], Y.shape[1], X.shape[0]
    x_loadings = np.zeros((B, n_components))
    y_loadings = np.zeros((T, n_components))
    x_scores = np.zeros((S, n_components))
    x_weights = np.zeros((B, n_components))
    basis = np.zeros((B, n_components))
    for comp in range(n_components):
        ci, si, ri = svd(Cov, n_components=1, seed=seed)
        ti = X0 @ ri
        nt = np.linalg.norm(ti)
        x_weights[:, [comp]] = ri / nt
        ti /= nt
        x_scores[:, [comp]] = ti
        x_loadings[:, [comp]] = X0.T @ ti
        y_loadings[:, [comp]] = Y0.T @ ti
        vi = x_loadings[:, [comp]]
        for _ in range(2):
            for j in range(comp):
                vj = basis[:, [j]]
                vi = vi - ((vj.T @ vi) * vj)
        vi /= np.linalg.norm(vi)
        basis[:, [comp]] = vi
        Cov = Cov - (vi @ (vi.T @ Cov))
        Vi = basis[:, :comp]
        Cov = Cov - (Vi @ (Vi.T @ Cov))
    pctvar = [np.sum(x_loadings ** 2, axis=0) / np.sum(X0 ** 2),
              np.sum(y_loadings ** 2, axis=0) / np.sum(Y0 ** 2)]
    return dict(x_weights=x_weights, x_loadings=x_loadings,
                y_loadings=y_loadings, x_scores=x_scores, pctvar=pctvar)


def resid_yscores(x_scores, y_scores):
    """Orthogonalise y_scores against preceding x_scores (double MGS).
    regression.py:9-45."""
    x_scores = np.array(x_scores)
    y_scores = np.array(y_scores, copy=True)
    for comp in range(x_scores.shape[1]):
        ui = y_scores[:, [comp]]
        for _ in range(2):
            for j in range(comp):
                tj = x_scores[:, [j]]
                ui = ui - ((tj.T @ ui) * tj)
        y_scores[:, [comp]] = ui
    return y_scores


def _regression_single_perm(spec, X, Y, inds, seed):
    """regression.py:329-373 -- y_weights is None in the caller so the rotate
    branch (:359) never runs; returns pctvar in Y per component."""
    x_weights, vexp, _ = decompose(spec, X, Y[inds], seed=seed)
    return np.diag(vexp)


def _regression_single_boot(spec, X, Y, inds, original, seed):
    """regression.py:279-327 (2-D Y): sign-align against the original weights
    through efficient_corr, y_loadings = Yi.T @ (Xi @ w)."""
    Xi, Yi = X[inds], Y[inds]
    x_weights = decompose(spec, Xi, Yi, seed=seed)[0]
    if original is not None:
        x_weights = x_weights * np.sign(efficient_corr(x_weights, original))
    mask = get_mask(Xi, Yi)
    y_loadings = Yi[mask].T @ (Xi @ x_weights)[mask]
    return y_loadings, x_weights


# --------------------------------------------------------------------------
# whole analyses (what the reference's front-end functions return, restricted
# to the keys the hot path feeds).  test_split / n_split are always off.
# --------------------------------------------------------------------------

# --------------------------------------------------------------------------
# vectorised "fast oracle" (SURVEY 7.1 step 1 / 8d): the statistics that do NOT
# need a B-sized product per resample -- permuted singular values (hence
# p-values) and the bootstrap distribution (hence CIs) -- at full n_perm /
# n_boot on the CPU in seconds.  It restates the SAME reference functions
# through two exact identities and is itself held to the per-resample oracle
# above (tests/test_oracle_golden.py::test_fast_oracle_*):
#   * X is fixed under a behavioural permutation and the z-score is per cell, so
#     R = A_pi Zx with Zx the cell-wise standardised X and A_pi (K, S) holding the
#     standardised permuted Y; the rotated singular values
#     sqrt(sum(procrustes(V, V_pi, d_pi)**2, 0)) (pyls/base.py:699-700) equal
#     |R^T v_j| = sqrt(v_j^T A (Zx Zx^T) A^T v_j) whenever K <= B, the un-rotated
#     ones are sqrt(eig(A (Zx Zx^T) A^T)).  Mean-centred: R = (C P_pi) X.
#   * gen_distrib only uses the ORIGINAL weights (pyls/base.py:574), and both
#     variants are linear in X before the projection:
#     gen_distrib(X[i], Y[i], U) = gen_distrib((X @ normalize(U))[i], Y[i], I).
# --------------------------------------------------------------------------

def _perm_operator(spec, Y, perminds):
    """(K, S) left operand A with R = A @ Xfixed for one permutation."""
    S = len(perminds)
    if spec.kind == 'meancentered':
        # R = C X[pi]: scatter the columns of the centring operator through pi
        C = getattr(spec, '_centring_operator', None)
        if C is None:
            C = spec._centring_operator = get_mean_center(
                np.eye(S), spec.dummy, spec.n_cond, spec.mean_centering,
                means=True)
        A = np.zeros_like(C)
        np.add.at(A.T, perminds, C.T)
        return A
    Yp = Y[perminds]
    T = Y.shape[1]
    cells = spec.dummy.T.astype(bool)
    A = np.zeros((len(cells) * T, S))
    for g, c in enumerate(cells):
        y = Yp[c]
        yn = y - y.mean(axis=0)
        if not spec.covariance:
            yn = yn / y.std(axis=0, ddof=1)
        A[g * T:(g + 1) * T, c] = yn.T / (c.sum() - 1)
    return A


def fixed_gram(spec, X):
    """S x S Gram matrix of the data matrix the permutations contract with."""
    if spec.kind == 'meancentered':
        return X @ X.T
    Zx = np.zeros_like(X, dtype=float)
    for c in spec.dummy.T.astype(bool):
        x = X[c]
        Zx[c] = x - x.mean(axis=0)
        if not spec.covariance:
            Zx[c] /= x.std(axis=0, ddof=1)
    return Zx @ Zx.T


def fast_perm_singvals(spec, X, Y, permsamp, original_v, gram=None):
    """d_perm (L, P) for index permutations, in sample space (see above).
    Needs K <= B (true for every benchmark configuration)."""
    gram = fixed_gram(spec, X) if gram is None else gram
    out = np.empty((original_v.shape[1], permsamp.shape[1]))
    for i in range(permsamp.shape[1]):
        A = _perm_operator(spec, Y, permsamp[:, i])
        if spec.rotate:
            W = original_v.T @ A
            out[:, i] = np.sqrt(np.maximum(np.sum((W @ gram) * W, axis=1), 0))
        else:
            lam = np.linalg.eigvalsh(A @ gram @ A.T)[::-1]
            out[:, i] = np.sqrt(np.maximum(lam, 0))[:out.shape[0]]
    return out


def fast_boot_distrib(spec, X, Y, bootsamp, original_u):
    """distrib (K, L, R): gen_distrib of every bootstrap evaluated on the
    projected scores X @ normalize(U_orig) (S x L) instead of on X."""
    Sx = X @ normalize(original_u)
    eye = np.eye(Sx.shape[1])
    out = [gen_distrib(spec, Sx[bootsamp[:, i]],
                       None if spec.kind == 'meancentered' else
                       Y[bootsamp[:, i]], eye)
           for i in range(bootsamp.shape[1])]
    return np.stack(out, axis=-1)


def fast_stats(spec, X, Y, permsamp, bootsamp, U, d, V, ci=95):
    """p-values and bootstrap percentile intervals at full n from the two
    functions above -> dict(pvals, perm_singval, distrib, distrib_ci)."""
    d_perm = fast_perm_singvals(spec, X, Y, permsamp, V)
    distrib = fast_boot_distrib(spec, X, Y, bootsamp, U)
    d = np.diag(d) if np.ndim(d) == 1 else d
    return dict(perm_singval=d_perm, pvals=perm_sig(d, d_perm),
                distrib=distrib,
                distrib_ci=np.stack(boot_ci(distrib, ci=ci), -1))


def _finish_boot(res, orig_bs, distrib, u_sum, u_square, n, ci, add_orig):
    if add_orig:
        u_sum, u_square = u_sum + orig_bs, u_square + orig_bs ** 2
    bsr, se = boot_rel(orig_bs, u_sum, u_square, n)
    res['x_weights_normed'] = bsr
    res['x_weights_stderr'] = se
    res['distrib'] = distrib
    res['distrib_ci'] = np.stack(boot_ci(distrib, ci=ci), -1)


def behavioral_pls(X, Y, groups=None, n_cond=1, n_perm=5000, n_boot=5000,
                   covariance=False, rotate=True, ci=95, permsamples=None,
                   bootsamples=None, seed=None, permindices=True,
                   test_split=0, test_size=0.25, n_split=None):
    """pyls.behavioral_pls.  Follows
    pyls/base.py:341-399 and pyls/types/behavioral.py:172-227, including the
    order in which the seeded RandomState is consumed (original SVD ->
    gen_permsamp -> gen_bootsamp).  permindices=False: ``permsamples`` is a
    (P, S, T) stack of pre-permuted Y matrices (base.py:636-639)."""
    X, Y = np.asarray(X), np.asarray(Y)
    groups = [len(X) // n_cond] if groups is None else list(np.atleast_1d(groups))
    spec = _Spec('behavioral', groups, n_cond, covariance=covariance,
                 rotate=rotate)
    rs = check_random_state(seed)
    res = {}
    U, d, V = decompose(spec, X, Y, seed=rs)
    res['x_weights'], res['singvals_diag'], res['y_weights'] = U, d, V
    res['x_scores'] = X @ U
    if n_perm > 0:
        if permsamples is None:
            permsamples = gen_permsamp(groups, n_cond, n_perm, seed=rs)
        table = permsamples if permindices else \
            np.transpose(permsamples, (1, 2, 0))
        if n_split:
            d_perm, ucorrs, vcorrs = run_perms_split(
                spec, X, Y, table, V, n_split, use_permind=permindices)
        else:
            d_perm = run_perms(spec, X, Y, table, V, use_permind=permindices)
        res['pvals'] = perm_sig(d, d_perm)
        res['permsamples'] = permsamples
        res['perm_singval'] = d_perm
        if n_split:
            _split_results(spec, X, Y, res, U, d, V, n_split, ucorrs, vcorrs,
                           rs, ci)
    cells = np.repeat(groups, n_cond)
    res['y_scores'] = np.vstack([
        y @ v for y, v in zip(np.split(Y, np.cumsum(cells)[:-1]),
                              np.split(V, len(cells)))])
    res['y_loadings'] = gen_covcorr(spec, res['x_scores'], Y)
    if n_boot > 0:
        if bootsamples is None:
            bootsamples = gen_bootsamp(groups, n_cond, n_boot, seed=rs)
        distrib, u_sum, u_square = run_boots(spec, X, Y, bootsamples, U)
        res['bootsamples'] = bootsamples
        _finish_boot(res, U @ d, distrib, u_sum, u_square, n_boot + 1, ci,
                     add_orig=True)
    if test_split and test_size > 0:
        # pyls/types/behavioral.py:217-219: seeded by the analysis' RandomState
        # after the bootstrap table has been drawn
        res['pearson_r'], res['r_squared'] = crossval(
            spec, X, Y, test_split, test_size=test_size, seed=rs)
    res['varexp'] = np.diag(varexp(d))
    res['singvals'] = np.diag(d)
    return res


def meancentered_pls(X, groups=None, n_cond=1, mean_centering=0, n_perm=5000,
                     n_boot=5000, rotate=True, ci=95, permsamples=None,
                     bootsamples=None, seed=None, n_split=None):
    """pyls.meancentered_pls with permindices=True.  Follows
    pyls/types/meancentered.py:11-48 (argument fix-ups) and :127-179."""
    X = np.asarray(X)
    groups = [len(X) // n_cond] if groups is None else list(np.atleast_1d(groups))
    if n_cond == 1 and len(groups) == 1:
        raise ValueError('Cannot perform PLS with only one group and one '
                         'condition. Please confirm inputs are correct.')
    if n_cond == 1 and mean_centering == 0:
        mean_centering = 1
    elif len(groups) == 1 and mean_centering == 1:
        mean_centering = 0
    spec = _Spec('meancentered', groups, n_cond,
                 mean_centering=mean_centering, rotate=rotate)
    Y = spec.dummy
    rs = check_random_state(seed)
    res = {}
    U, d, V = decompose(spec, X, Y, seed=rs)
    res['x_weights'], res['singvals_diag'], res['y_weights'] = U, d, V
    res['x_scores'] = X @ U
    if n_perm > 0:
        if permsamples is None:
            permsamples = gen_permsamp(groups, n_cond, n_perm, seed=rs)
        if n_split:
            d_perm, ucorrs, vcorrs = run_perms_split(spec, X, Y, permsamples,
                                                     V, n_split)
        else:
            d_perm = run_perms(spec, X, Y, permsamples, V)
        res['pvals'] = perm_sig(d, d_perm)
        res['permsamples'] = permsamples
        res['perm_singval'] = d_perm
        if n_split:
            _split_results(spec, X, Y, res, U, d, V, n_split, ucorrs, vcorrs,
                           rs, ci)
    res['y_scores'] = Y @ V
    dm = get_mean_center(X, Y, n_cond, mean_centering, False) @ U
    res['contrast'] = np.vstack([dm[c].mean(axis=0)
                                 for c in Y.T.astype(bool)])
    if n_boot > 0:
        if bootsamples is None:
            bootsamples = gen_bootsamp(groups, n_cond, n_boot, seed=rs)
        distrib, u_sum, u_square = run_boots(spec, X, Y, bootsamples, U)
        res['bootsamples'] = bootsamples
        _finish_boot(res, U @ d, distrib, u_sum, u_square, n_boot, ci,
                     add_orig=False)
    res['varexp'] = np.diag(varexp(d))
    res['singvals'] = np.diag(d)
    return res


AGGFUNCS = dict(mean=np.mean, median=np.median, sum=np.sum)


def pls_regression(X, Y, n_components=None, n_perm=5000, n_boot=5000,
                   rotate=True, ci=95, permsamples=None, bootsamples=None,
                   seed=None, aggfunc='mean'):
    """pyls.pls_regression.  Follows pyls/types/regression.py:190-246 and
    :375-428.  Unlike the reference the caller's arrays are not centred in
    place (the oracle works on copies).

    Three-dimensional Y (S, T, C) (regression.py:207-235, 308-310, 391-399):
    the decomposition and the permutations run on ``aggfunc(Y, axis=-1)``;
    bootstrap i draws rows ``s[:, i]`` AND a bootstrap sample ``c[:, i]`` of
    the third axis, aggregating a fresh, un-centred behaviour matrix
    ``aggfunc(Y[..., c[:, i]], axis=-1)[s[:, i]]``.  `bootsamples` is then the
    pair of tables ``(s (S, n_boot), c (C, n_boot))``."""
    X, Y = np.array(X, dtype=float), np.array(Y, dtype=float)
    max_comp = min(len(X) - 1, X.shape[1])
    n_components = max_comp if n_components is None else int(n_components)
    if n_components > max_comp:
        raise ValueError('Provided `n_components` cannot be greater '
                         'than {}'.format(max_comp))
    groups = [len(X)]
    spec = _Spec('regression', groups, 1, rotate=rotate,
                 n_components=n_components)
    rs = check_random_state(seed)
    Y3 = None
    if Y.ndim == 3:
        agg = AGGFUNCS.get(aggfunc, aggfunc)
        Y3, Y = Y, agg(Y, axis=-1)
        if n_boot > 0 and bootsamples is None:
            bootsamples = (gen_bootsamp([Y3.shape[0]], 1, n_boot, seed=seed),
                           gen_bootsamp([Y3.shape[-1]], 1, n_boot, seed=seed))
    X -= np.nanmean(X, axis=0, keepdims=True)
    Y -= np.nanmean(Y, axis=0, keepdims=True)
    mask = get_mask(X, Y)
    res = {}
    W, vexp, _ = decompose(spec, X, Y, seed=rs)
    res['x_weights'] = W
    res['x_scores'] = X @ W
    if n_perm > 0:
        if permsamples is None:
            permsamples = gen_permsamp(groups, 1, n_perm, seed=rs)
        d_perm = run_perms(spec, X, Y, permsamples, None)
        res['pvals'] = perm_sig(vexp, d_perm)
        res['permsamples'] = permsamples
        res['perm_singval'] = d_perm
    res['y_loadings'] = Y[mask].T @ res['x_scores'][mask]
    res['y_scores'] = np.full((len(Y), n_components), np.nan)
    res['y_scores'][mask] = resid_yscores(res['x_scores'][mask],
                                          Y[mask] @ res['y_loadings'])
    if n_boot > 0 and Y3 is not None:
        srows, third = bootsamples
        u_sum, u_square, distrib = np.zeros_like(W), np.zeros_like(W), []
        for i in range(n_boot):
            # regression.py:308-310: the aggregated matrix is used as it is
            Yi = agg(Y3[..., third[:, i]], axis=-1)
            d, u = _regression_single_boot(spec, X, Yi, srows[:, i], W, i)
            u_sum += u
            u_square += u ** 2
            distrib.append(d)
        res['bootsamples'] = bootsamples
        _finish_boot(res, W, np.stack(distrib, axis=-1), u_sum, u_square,
                     n_boot + 1, ci, add_orig=True)
    elif n_boot > 0:
        if bootsamples is None:
            bootsamples = gen_bootsamp(groups, 1, n_boot, seed=rs)
        distrib, u_sum, u_square = run_boots(spec, X, Y, bootsamples, W)
        res['bootsamples'] = bootsamples
        _finish_boot(res, W, distrib, u_sum, u_square, n_boot + 1, ci,
                     add_orig=True)
    res['varexp'] = np.diag(vexp)
    return res
