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
from . import messages, structures
from .engine import Download, ResamplingEngine, copy_stream, to_host
from .resample import (check_bootsamples, check_random_state, gen_bootsamp,
                       gen_permsamp, gen_splits)


def _device_seed(rs):
    """64-bit key for the on-device generator, drawn from the analysis'
    RandomState so that an integer ``seed`` makes runs repeatable."""
    hi, lo = rs.randint(0, 2 ** 31 - 1, size=2)
    return (int(hi) << 32) | int(lo)


class _DeviceTable:
    """A device-generated resampling table, (n, S) int32, on its way to the
    host.  The copy (side stream) is queued as soon as the table exists --
    transposed and widened on the device to the (S, n) int64 array the
    reference keeps, so the host does no conversion -- and ``get()`` waits for
    that copy only."""

    def __init__(self, full):
        self.dl = Download(full.t().contiguous().to(torch.int64))
        self.arr = None

    def start(self):
        """Kept for callers that used to trigger the copy themselves."""

    def get(self):
        if self.arr is None:
            self.arr = self.dl.get()
            self.dl = None
        return self.arr


def _resolve(table):
    return table.get() if isinstance(table, _DeviceTable) else table


def _host_t(t):
    return None if t is None else to_host(t).T.copy()


def _start(t):
    return None if t is None else Download(t)


def normalise_groups(groups, n_rows, n_cond):
    """`groups` as a list of ints: one group of every subject when it is not
    given, a scalar wrapped (the tolerance of pyls/base.py:255-262)."""
    if groups is None:
        return [n_rows // n_cond]
    if isinstance(groups, (list, tuple, np.ndarray)):
        return [int(g) for g in groups]
    return [int(groups)]


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
        Gram matrix of the data -- same values, no B-sized work) and
        ``gather_results`` (multi-process runs only: 'all' -- every rank
        returns the complete results, default; 'root' -- the per-resample
        arrays (bootstrap distribution, tables) and the B-sized arrays are
        moved to the host of rank 0 only, the other ranks return the
        statistics (p-values, intervals) and ``None`` for those arrays),
        ``input_source`` (multi-process runs only: 'local', default -- every rank
        uploads the arrays it was given; 'root' -- rank 0 uploads its arrays and
        the other ranks receive them over NCCL, their own copies only give the
        shapes), ``gemm_backend`` ('auto', default: the cross-covariance contraction
        runs as int8 digit-plane products on the tcgen05 tensor cores where
        that kernel applies; 'dmma': FP64 DMMA everywhere) and ``gemm_slices``
        (digit planes per operand, 5 / 6 / 7; default 6: entries within 1e-13
        of the FP64 product relative to the norms they contract).
    """

    engine_mode = None

    def __init__(self, X, Y=None, groups=None, n_cond=1, **kwargs):
        groups = normalise_groups(groups, len(X), n_cond)
        expected = sum(groups) * n_cond
        if expected != len(X):
            raise ValueError(messages.SAMPLES_MISMATCH.format(
                len(X), expected, groups, n_cond))
        if Y is not None and len(Y) != len(X):
            raise ValueError(messages.XY_ROWS_DIFFER.format(len(X), len(Y)))

        self.inputs = structures.PLSInputs(X=X, Y=Y, groups=groups,
                                           n_cond=n_cond, **kwargs)
        seed = self.inputs.get('seed')
        if pdist.world()[1] > 1 and \
                not isinstance(seed, (int, np.integer)):
            # every rank must draw the same tables, masks and device keys:
            # rank 0 fixes an integer seed for all of them
            seed = pdist.broadcast_seed(check_random_state(seed))
        self.rs = check_random_state(seed)
        mode = self.inputs.get('gather_results')
        if mode is None:
            self.inputs['gather_results'] = mode = 'all'
        if mode not in ('all', 'root'):
            raise ValueError("gather_results must be 'all' or 'root'")
        backend = self.inputs.get('index_backend')
        if backend is None:
            self.inputs['index_backend'] = backend = 'device'
        if backend not in ('device', 'reference'):
            raise ValueError("index_backend must be 'device' or 'reference'")
        if self.inputs.get('input_source') not in (None, 'local', 'root'):
            raise ValueError("input_source must be 'local' or 'root'")
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
                                   'workspace_bytes'),
                               gemm_backend=self.inputs.get(
                                   'gemm_backend') or 'auto',
                               gemm_slices=self.inputs.get(
                                   'gemm_slices') or 6)
        if self.inputs.get('input_source') == 'root':
            # multi-process runs: only rank 0 uploads, the replicas travel over NCCL
            X = pdist.broadcast_from_root(eng, X, (S, B))
            if Y is not None:
                Y = pdist.broadcast_from_root(eng, Y, (S, T))
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
        L = self.engine.L          # min(K, B): the short side of the matrix
        self.rs.normal(size=(L, L + 10))

    # -- analysis ------------------------------------------------------------
    def run_pls(self, X, Y):
        """
        Original decomposition + permutation test
        (pyls/base.py:341-399).  Only QUEUES the device work and registers the
        downloads in ``self._later``: the results object is complete after the
        caller's ``self._finalize()``.  Device copies of the decomposition stay
        in ``self._dev``.
        """
        self.res = res = structures.PLSResults(inputs=self.inputs)
        self._later = []
        X = np.asarray(X) if not isinstance(X, torch.Tensor) else X
        self.engine = eng = self._make_engine(X, self._engine_y(Y))
        U, d, V = eng.decompose()
        self._replay_svd_draws()
        self._dev = dict(U=U, d=d, V=V)
        dl = [Download(t) for t in (U, d, V, eng.project_scores(U))]

        def fill_decomposition():
            res['x_weights'] = dl[0].get()
            res['singvals'] = np.diag(dl[1].get())
            res['y_weights'] = dl[2].get()
            res['x_scores'] = dl[3].get()
        self._later.append(fill_decomposition)

        if self.inputs.n_perm > 0:
            out = self._permutation_device(X, Y, seed=self.rs)
            pv = eng.perm_pvals(out['d_perm'], d)
            stats = {}
            if self.inputs.n_split is not None:
                # split-half reliability of the original singular vectors and
                # its permutation statistics (pyls/base.py:373-397)
                ou, ov = self.split_half(X, Y, seed=self.rs)
                ci = self.inputs.get('ci')
                ci = 95 if ci is None else ci
                low = (100 - ci) / 2
                for name, orig, perm in (('ucorr', ou, out['ucorrs']),
                                         ('vcorr', ov, out['vcorrs'])):
                    lo, hi = eng.percentile(perm, low, 100 - low)
                    stats[name] = orig
                    stats[name + '_pvals'] = eng.perm_pvals(perm, orig)
                    stats[name + '_lolim'] = lo
                    stats[name + '_uplim'] = hi
            dl_pv, dl_dperm = Download(pv), Download(out['d_perm'])
            dl_stats = {k: Download(v) for k, v in stats.items()}

            def fill():
                res['permres']['pvals'] = dl_pv.get()
                self.permsamp = _resolve(out['table'])
                res['permres']['permsamples'] = self.permsamp
                res['permres']['perm_singval'] = dl_dperm.get().T.copy()
                if dl_stats:
                    res['splitres'].update({k: v.get()
                                            for k, v in dl_stats.items()})
            self._later.append(fill)
        return res

    def _finalize(self):
        """Everything above only queues device work and starts the downloads
        (side stream) behind their producers; the host conversions registered
        on the way run here in queue order, each waiting for its own download
        only -- the early ones while the later kernels are still running."""
        later, self._later = self._later, []
        for fill in later:
            fill()
        if self.engine is not None:
            torch.cuda.current_stream(self.engine.device).synchronize()

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
        host, or generated on the device.  Returns (table, block, first):
        `block` is the device (n_local, S) int32 block of this rank starting at
        resample id `first`; `table` is the (S, n) int array of the whole
        table or, for device-generated tables, a :class:`_DeviceTable` whose
        copy to the host the caller starts after queuing its kernels."""
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
            # behavioural analyses with several cells contract every cell's operand
            # rows over that cell's rows of X only
            if kind == 'boot' and eng.mode.startswith('behavioral') and \
                    eng.J > 1 and given.shape[0] == eng.S and \
                    given.size and 0 <= given.min() and given.max() < eng.S:
                check_bootsamples(given, self.inputs.groups,
                                  self.inputs.n_cond)
            block = eng.to_device_indices(given[:, first:first + count])
            return given, block, first
        gen = eng.gen_perm_indices if kind == 'perm' else eng.gen_boot_indices
        block, exhausted = gen(_device_seed(check_random_state(seed)), count,
                               first=first)
        if exhausted:
            warnings.warn('WARNING: Duplicate {} used.'.format(
                'permutations' if kind == 'perm' else 'bootstraps'))
        full = pdist.gather_resamples(block, n)
        return _DeviceTable(full), block, first

    def permutation(self, X, Y, seed=None):
        """
        Permutation test on the device (replaces pyls/base.py:601-712).

        Returns
        -------
        d_perm : (L, P) numpy.ndarray
        ucorrs, vcorrs : (L, P) numpy.ndarray or None
            Split-half correlations of every permutation (``n_split``)
        """
        out = self._permutation_device(X, Y, seed=seed)
        self.permsamp = _resolve(out['table'])
        return (_host_t(out['d_perm']), _host_t(out['ucorrs']),
                _host_t(out['vcorrs']))

    def _permutation_device(self, X, Y, seed=None):
        """Queues the permutation test; returns the device results
        (``d_perm`` (P, L), ``ucorrs`` / ``vcorrs`` (P, L) or None) and the
        table (array or :class:`_DeviceTable`) without waiting for anything."""
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
            given = np.asarray(given)
            local = self._prepermuted(given, n, rotate)
            # the reference keeps the stack transposed, (S, T, P)
            # (pyls/base.py:638-639, 370)
            table = np.transpose(given, (1, 2, 0))
            first, count = pdist.my_block(n)
            split_of = dict(Yperm=given[first:first + count])
        else:
            path = self.inputs.get('perm_path') or 'gemm'
            if path not in ('gemm', 'gram'):
                raise ValueError("perm_path must be 'gemm' or 'gram'")
            table, block, first = self._table('perm', n, seed)
            count = int(block.shape[0])
            if path == 'gram' and rotate:
                local = self.engine.run_perms_gram(block)
            else:
                local = self.engine.run_perms(block, rotate=rotate)
            if isinstance(table, _DeviceTable):
                table.start()            # behind the kernels queued above
            split_of = dict(idx=block)
        d_perm = pdist.gather_resamples(local, n)
        self._dev['d_perm'] = d_perm
        out = dict(d_perm=d_perm, ucorrs=None, vcorrs=None, table=table)
        if n_split is None:
            return out
        # split-half resampling of every permuted data set (pyls/base.py:704-708)
        uc, vc = self.engine.split_half(self._split_masks(first, count),
                                        **split_of)
        uc, vc = pdist.gather_resamples(uc, n), pdist.gather_resamples(vc, n)
        self._dev.update(ucorrs=uc, vcorrs=vc)
        out.update(ucorrs=uc, vcorrs=vc)
        return out

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
        out = self._bootstrap_device(X, Y, seed=seed)
        us, uq = to_host(out['u_sum']), to_host(out['u_square'])   # waits
        self.bootsamp = _resolve(out['table'])
        return self._host_distrib(out), us, uq

    @staticmethod
    def _host_distrib(out):
        """(K, L, R) host view of the bootstrap distribution of a
        :meth:`_bootstrap_device`; waits for its copies."""
        if out['host'] is not None:
            out['host_done'].synchronize()
            return out['host'].numpy().transpose(1, 2, 0)
        if out['distrib'] is None:
            return None                  # gather_results='root', other ranks
        return to_host(out['distrib'].permute(1, 2, 0).contiguous())

    def _bootstrap_device(self, X, Y, seed=None):
        """Queues the bootstrap; returns device ``distrib`` (R, K, L), ``u_sum``,
        ``u_square`` (B, L), the pinned host copy of ``distrib`` that is being
        filled on a side stream (``host_done`` marks its last copy) and the
        table, without waiting."""
        n = self.inputs.n_boot
        table, block, first = self._table('boot', n, seed)
        count = int(block.shape[0])
        rank, size = pdist.world()
        root_only = size > 1 and self.inputs.gather_results == 'root'
        # the bootstrap distribution needs the original weights only: it is computed
        # first and moves to pinned host memory (side stream) under the
        # cross-covariance work queued behind it.  Several ranks: the blocks are
        # all-gathered on the device (or gathered to rank 0 alone,
        # gather_results='root') and every host copies the result ONCE
        want_host = not root_only or rank == 0
        host = torch.empty((n, self.engine.K, self.engine.L),
                           dtype=torch.float64, pin_memory=True) \
            if want_host else None
        distrib = local = self.engine.boot_distrib(block)
        main = torch.cuda.current_stream(self.engine.device)
        side = copy_stream(self.engine.device)
        ready = torch.cuda.Event()
        ready.record(main)
        host_done = torch.cuda.Event()
        with torch.cuda.stream(side):
            side.wait_event(ready)
            if size > 1:
                distrib = pdist.gather_to_root(local, n) if root_only \
                    else pdist.gather_resamples(local, n)
            if want_host:
                host.copy_(distrib, non_blocking=True)
            host_done.record(side)
        if distrib is not None:
            distrib.record_stream(side)
        _, u_sum, u_square = self.engine.run_boots(block, want_distrib=False)
        if size > 1:
            pdist.reduce_sum(u_sum, u_square)
        if isinstance(table, _DeviceTable):
            table.start()                # behind the kernels queued above
        self._dev.update(distrib=distrib, distrib_local=local, u_sum=u_sum,
                         u_square=u_square)
        return dict(distrib=distrib, host=host, host_done=host_done,
                    u_sum=u_sum, u_square=u_square, table=table, keep=local)

    def _boot_stats(self, add_orig, device=False):
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
        # several ranks: every rank selects the order statistics of its share of
        # the K x L series over all resamples (one all-to-all), not of everything
        lo, hi = pdist.percentile_sharded(eng, dev['distrib_local'],
                                          self.inputs.n_boot, low, 100 - low)
        out = (bsr, se, torch.stack([lo, hi], dim=-1))
        return out if device else tuple(to_host(t) for t in out)
