# -*- coding: utf-8 -*-
"""
Resampling driver with the interface of the reference's ``BasePLS``
(pyls/base.py:232-770): ``run_pls`` -> ``permutation()`` / ``bootstrap()``.

Where the reference maps ``_single_perm`` / ``_single_boot`` over resamples in
Python (optionally in joblib worker processes), this class hands the whole
batch to the CUDA engine: ONE host -> C-ABI -> device crossing per analysis
step.  With ``torch.distributed`` initialised (one process per GPU) every rank
runs a contiguous block of resample ids and the results are all-gathered /
all-reduced (see :mod:`pypyls_b200.dist`).
"""

import warnings

import numpy as np
import torch

from . import dist as pdist
from . import structures
from .engine import ResamplingEngine, to_host
from .resample import (check_random_state, gen_bootsamp, gen_permsamp,
                       gen_splits)


def _device_seed(rs):
    """64-bit key for the on-device generator, drawn from the analysis'
    RandomState so that an integer ``seed`` makes runs repeatable."""
    hi, lo = rs.randint(0, 2 ** 31 - 1, size=2)
    return (int(hi) << 32) | int(lo)


class BasePLS():
    """
    Base class of the PLS types.

    Parameters
    ----------
    X : (S, B) array_like
    Y : (S, T) array_like, optional
    groups : (G,) list of int, optional
    n_cond : int, optional
    **kwargs : see :obj:`pypyls_b200.structures.PLSInputs`.  Additions to the
        reference's keys: ``index_backend`` ('device' -- counter-based
        generator on the GPU, default; 'reference' -- host tables replaying
        the reference's NumPy stream for the given seed), ``device`` (CUDA
        ordinal), ``workspace_bytes`` and ``perm_path`` ('gemm': every
        permutation runs the cross-covariance contraction, default; 'gram':
        rotated permutations are evaluated in sample space through the S x S
        Gram matrix of the data -- same values, no B-sized work).
    """

    engine_mode = None

    def __init__(self, X, Y=None, groups=None, n_cond=1, **kwargs):
        if groups is None:
            groups = [len(X) // n_cond]
        elif not isinstance(groups, (list, np.ndarray)):
            groups = [groups]
        groups = [int(g) for g in groups]

        n_samples = sum([g * n_cond for g in groups])
        if len(X) != n_samples:
            raise ValueError('Number of samples specified by `groups` and '
                             '`n_cond` does not match number of samples in '
                             'input array(s).\n'
                             '    EXPECTED: {}\n'
                             '    ACTUAL:   {} (groups: {} * n_cond: {})'
                             .format(len(X), n_samples, groups, n_cond))
        if Y is not None and len(X) != len(Y):
            raise ValueError('Provided `X` and `Y` matrices must have the '
                             'same number of samples. Provided matrices '
                             'differed: X: {}, Y: {}'.format(len(X), len(Y)))

        self.inputs = structures.PLSInputs(X=X, Y=Y, groups=groups,
                                           n_cond=n_cond, **kwargs)
        self.rs = check_random_state(self.inputs.get('seed'))
        backend = self.inputs.get('index_backend')
        if backend is None:
            self.inputs['index_backend'] = backend = 'device'
        if backend not in ('device', 'reference'):
            raise ValueError("index_backend must be 'device' or 'reference'")
        self.engine = None

    # -- engine ------------------------------------------------------------
    def _make_engine(self, X, Y):
        S, B = X.shape
        T = 1 if Y is None else Y.shape[1]
        device = self.inputs.get('device')
        if device is None:
            device = torch.cuda.current_device() \
                if torch.cuda.is_available() else 0
        eng = ResamplingEngine(self.engine_mode(), S, B, T,
                               self.inputs.groups, self.inputs.n_cond,
                               mean_centering=self.inputs.get(
                                   'mean_centering') or 0,
                               device=device,
                               workspace_bytes=self.inputs.get(
                                   'workspace_bytes'))
        eng.set_data(X, Y)
        return eng

    def engine_mode(self):
        raise NotImplementedError

    def _engine_y(self, Y):
        return Y

    def _replay_svd_draws(self):
        """The reference's original decomposition draws a (K, K+10) normal
        matrix from the analysis' RandomState inside sklearn's randomized_svd
        (pyls/base.py:362-363 -> pyls/compute.py:44-45) before any resampling
        table is generated; consume the same draws so that
        index_backend='reference' reproduces the reference's tables."""
        K = self.engine.K
        self.rs.normal(size=(K, K + 10))

    # -- analysis ------------------------------------------------------------
    def run_pls(self, X, Y):
        """
        Original decomposition + permutation test
        (pyls/base.py:341-399).  Returns the results object; device copies of
        the decomposition stay in ``self._dev``.
        """
        self.res = res = structures.PLSResults(inputs=self.inputs)
        X = np.asarray(X) if not isinstance(X, torch.Tensor) else X
        self.engine = eng = self._make_engine(X, self._engine_y(Y))
        U, d, V = eng.decompose()
        self._replay_svd_draws()
        self._dev = dict(U=U, d=d, V=V)
        res['x_weights'] = to_host(U)
        res['singvals'] = np.diag(to_host(d))
        res['y_weights'] = to_host(V)
        res['x_scores'] = to_host(eng.project_scores(U))

        if self.inputs.n_perm > 0:
            d_perm, ucorrs, vcorrs = self.permutation(X, Y, seed=self.rs)
            res['permres']['pvals'] = eng.perm_pvals(
                self._dev['d_perm'], d).cpu().numpy()
            res['permres']['permsamples'] = self.permsamp
            res['permres']['perm_singval'] = d_perm

            if self.inputs.n_split is not None:
                # split-half reliability of the original singular vectors and
                # its permutation statistics (pyls/base.py:373-397)
                ou, ov = self.split_half(X, Y, seed=self.rs)
                dev = self._dev
                ci = self.inputs.get('ci')
                ci = 95 if ci is None else ci
                low = (100 - ci) / 2
                out = {}
                for name, orig, perm in (('ucorr', ou, dev['ucorrs']),
                                         ('vcorr', ov, dev['vcorrs'])):
                    lo, hi = eng.percentile(perm, low, 100 - low)
                    out[name] = to_host(orig)
                    out[name + '_pvals'] = to_host(eng.perm_pvals(perm, orig))
                    out[name + '_lolim'] = to_host(lo)
                    out[name + '_uplim'] = to_host(hi)
                res['splitres'].update(out)
        return res

    def _split_masks(self, first, count):
        """Half / half masks of the permutations [first, first + count):
        (count, n_split, S).  The reference draws a fresh set inside every
        permutation from ``RandomState(i)``, i the permutation's number
        (pyls/base.py:646-648, 704-708); index_backend='reference' replays
        that on the host, 'device' generates them with the counter-based
        generator on the GPU."""
        n_split = self.inputs.n_split
        if self.inputs.index_backend == 'reference':
            return np.stack([
                gen_splits(self.inputs.groups, self.inputs.n_cond, n_split,
                           seed=i, test_size=0.5).T
                for i in range(first, first + count)]).astype(np.int32)
        masks, exhausted = self.engine.gen_split_masks(
            self._split_seed, count, n_split, first=first)
        if exhausted:
            warnings.warn('WARNING: Duplicate split halves used.')
        return masks

    def split_half(self, X, Y, seed=None):
        """
        Split-half correlations of the original decomposition on the device
        (replaces pyls/base.py:714-770 as called from run_pls, :373-380).

        Returns
        -------
        ucorr, vcorr : (L,) device tensors
        """
        n_split = self.inputs.n_split
        if self.inputs.index_backend == 'reference':
            masks = gen_splits(self.inputs.groups, self.inputs.n_cond, n_split,
                               seed=seed, test_size=0.5).T[None]
            masks = masks.astype(np.int32)
        else:
            masks, exhausted = self.engine.gen_split_masks(
                _device_seed(check_random_state(seed)), 1, n_split)
            if exhausted:
                warnings.warn('WARNING: Duplicate split halves used.')
        uc, vc = self.engine.split_half(masks, use_original=True)
        return uc[0], vc[0]

    def _table(self, kind, n, seed):
        """Resampling table for this analysis: user-provided, replayed on the
        host, or generated on the device.  Returns (host, block, first):
        `block` is the device (n_local, S) int32 block of this rank starting at
        resample id `first`; `host` is a callable that gives the (S, n) int
        array of the whole table -- for device-generated tables the copy to
        the host is deferred so that it overlaps the resampling kernels the
        caller queues first."""
        eng = self.engine
        key = 'permsamples' if kind == 'perm' else 'bootsamples'
        first, count = pdist.my_block(n)
        given = self.inputs.get(key)
        if given is None and self.inputs.index_backend == 'reference':
            gen = gen_permsamp if kind == 'perm' else gen_bootsamp
            given = gen(self.inputs.groups, self.inputs.n_cond, n, seed=seed,
                        verbose=self.inputs.get('verbose'))
        if given is not None:
            given = np.asarray(given)
            if given.ndim != 2 or given.shape[-1] != n:
                raise ValueError('Provided `{}` must have shape ({}, {}); got '
                                 '{}'.format(key, eng.S, n, given.shape))
            block = eng.to_device_indices(given[:, first:first + count])
            return (lambda: given), block, first
        gen = eng.gen_perm_indices if kind == 'perm' else eng.gen_boot_indices
        block, exhausted = gen(_device_seed(check_random_state(seed)), count,
                               first=first)
        if exhausted:
            warnings.warn('WARNING: Duplicate {} used.'.format(
                'permutations' if kind == 'perm' else 'bootstraps'))
        full = pdist.gather_resamples(block, n)
        return (lambda: to_host(full).T.astype(int)), block, first

    def permutation(self, X, Y, seed=None):
        """
        Permutation test on the device (replaces pyls/base.py:601-712).

        Returns
        -------
        d_perm : (L, P) numpy.ndarray
        ucorrs, vcorrs : (L, P) numpy.ndarray or None
            Split-half correlations of every permutation (``n_split``)
        """
        n = self.inputs.n_perm
        rotate = self.inputs.get('rotate')
        rotate = True if rotate is None else bool(rotate)
        given = self.inputs.get('permsamples')
        n_split = self.inputs.n_split
        if n_split is not None and self.inputs.index_backend != 'reference':
            self._split_seed = _device_seed(check_random_state(seed))
        split_of = None
        if given is not None and self.inputs.get('permindices') is False:
            # pre-permuted Y matrices, (P, S, T) (pyls/base.py:636-639, 689-692)
            local = self._prepermuted(given, n, rotate)
            first, count = pdist.my_block(n)
            split_of = dict(Yperm=given[first:first + count])
        else:
            path = self.inputs.get('perm_path') or 'gemm'
            if path not in ('gemm', 'gram'):
                raise ValueError("perm_path must be 'gemm' or 'gram'")
            host_table, block, first = self._table('perm', n, seed)
            count = int(block.shape[0])
            if path == 'gram' and rotate:
                local = self.engine.run_perms_gram(block)
            else:
                local = self.engine.run_perms(block, rotate=rotate)
            self.permsamp = host_table()     # overlaps the kernels queued above
            split_of = dict(idx=block)
        d_perm = pdist.gather_resamples(local, n)
        self._dev['d_perm'] = d_perm
        if n_split is None:
            return to_host(d_perm).T.copy(), None, None
        # split-half resampling of every permuted data set (pyls/base.py:704-708)
        uc, vc = self.engine.split_half(self._split_masks(first, count),
                                        **split_of)
        uc, vc = pdist.gather_resamples(uc, n), pdist.gather_resamples(vc, n)
        self._dev.update(ucorrs=uc, vcorrs=vc)
        return (to_host(d_perm).T.copy(), to_host(uc).T.copy(),
                to_host(vc).T.copy())

    def _prepermuted(self, given, n, rotate):
        eng = self.engine
        if self.inputs.get('Y') is None or eng.T < 1 or \
                not self.engine_mode().startswith('behavioral'):
            raise ValueError('Pre-permuted `permsamples` (permindices=False) '
                             'need an analysis with a Y matrix.')
        shape = tuple(given.shape)
        if len(shape) != 3 or shape != (n, eng.S, eng.T):
            raise ValueError('Provided pre-permuted `permsamples` must have '
                             'shape ({}, {}, {}); got {}'.format(
                                 n, eng.S, eng.T, shape))
        self.permsamp = given
        first, count = pdist.my_block(n)
        return eng.run_perms_prepermuted(given[first:first + count],
                                         rotate=rotate)

    def bootstrap(self, X, Y, seed=None):
        """
        Bootstrap resampling on the device (replaces pyls/base.py:439-576).

        Returns
        -------
        distrib : (K, L, R) numpy.ndarray
        u_sum, u_square : (B, L) numpy.ndarray
        """
        n = self.inputs.n_boot
        host_table, block, _ = self._table('boot', n, seed)
        if pdist.world()[1] == 1:
            # single GPU: every internal pass's slice of `distrib` goes to the host
            # on a side stream while the next pass computes
            distrib, host, u_sum, u_square = \
                self.engine.run_boots_streamed(block)
            self.bootsamp = host_table()
            self._dev.update(distrib=distrib, u_sum=u_sum, u_square=u_square)
            us, uq = to_host(u_sum), to_host(u_square)   # synchronises the stream
            return host.numpy().transpose(1, 2, 0), us, uq
        distrib, u_sum, u_square = self.engine.run_boots(block)
        self.bootsamp = host_table()         # overlaps the kernels queued above
        distrib = pdist.gather_resamples(distrib, n)
        pdist.reduce_sum(u_sum, u_square)
        self._dev.update(distrib=distrib, u_sum=u_sum, u_square=u_square)
        return (to_host(distrib.permute(1, 2, 0).contiguous()),
                to_host(u_sum), to_host(u_square))

    def _boot_stats(self, add_orig):
        """Bootstrap ratios, standard errors and percentile intervals from the
        device-resident accumulators (compute.boot_rel / boot_ci,
        pyls/compute.py:184-237)."""
        eng, dev = self.engine, self._dev
        bs = dev['U'] * dev['d'][None, :]
        bsr, se = eng.boot_ratio(bs, dev['u_sum'], dev['u_square'],
                                 self.inputs.n_boot, add_orig)
        ci = self.inputs.get('ci')
        ci = 95 if ci is None else ci
        low = (100 - ci) / 2
        lo, hi = eng.percentile(dev['distrib'], low, 100 - low)
        return (to_host(bsr), to_host(se),
                to_host(torch.stack([lo, hi], dim=-1)))
