# -*- coding: utf-8 -*-
"""
Thin object wrapper over the C ABI of ``libplsb200.so``.

PyTorch is used for device memory and streams only: every tensor handed to the
library is a plain device pointer, all arithmetic happens in the hand-written
sm_100a kernels of ``csrc/``.  There is no CPU path -- constructing an engine
without a CUDA device or without the compiled library raises.
"""

import ctypes as C

import numpy as np
import torch

from . import _cabi

DEFAULT_WORKSPACE = 32 << 30     # bytes per pass of stored cross-covariances, capped at 1/5
                                 # of the device memory: B200 has 180 GB, and fewer, larger
                                 # passes keep every kernel's grid full

MODES = {
    'behavioral': _cabi.PLSB_BEHAVIORAL_CORR,
    'behavioral_cov': _cabi.PLSB_BEHAVIORAL_COV,
    'meancentered': _cabi.PLSB_MEANCENTERED,
    'regression': _cabi.PLSB_SIMPLS,
}


def to_host(t):
    """Device tensor -> NumPy array through pinned host memory (torch caches
    pinned blocks, so repeated analyses reuse them)."""
    if not t.is_cuda:
        return t.numpy()
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t)
    return out.numpy()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# Library handles (and the device workspaces they own) are pooled per device:
# an engine borrows a free handle and gives it back when closed, so repeated
# analyses do not pay cudaMalloc / cudaFree of the workspaces again.
_FREE_HANDLES = {}
_COPY_STREAMS = {}     # device index -> side stream for overlapped device->host copies


def copy_stream(device):
    """The side stream device -> host copies are queued on, so that they do
    not hold up the kernels queued behind them on the main stream."""
    index = torch.device(device).index
    side = _COPY_STREAMS.get(index)
    if side is None:
        side = _COPY_STREAMS[index] = torch.cuda.Stream(torch.device(device))
    return side


class Download:
    """A device tensor on its way to pinned host memory.  Creating it queues
    the copy on the side stream behind the work queued so far on the current
    stream; nothing blocks the host until :meth:`get`, which waits for this
    copy only and returns the NumPy array."""

    def __init__(self, t):
        self.dev = t                      # keeps the device memory alive
        self.host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        main, side = torch.cuda.current_stream(t.device), copy_stream(t.device)
        ready = torch.cuda.Event()
        ready.record(main)
        self.done = torch.cuda.Event()
        with torch.cuda.stream(side):
            side.wait_event(ready)
            self.host.copy_(t, non_blocking=True)
            self.done.record(side)

    def get(self):
        self.done.synchronize()
        self.dev = None
        return self.host.numpy()


def release_workspaces():
    """Destroys every pooled library handle (frees their device memory)."""
    lib = _cabi.lib()
    for handles in _FREE_HANDLES.values():
        for h, idle in handles:
            idle.synchronize()
            lib.plsb_destroy(h)
    _FREE_HANDLES.clear()


class ResamplingEngine:
    """
    One analysis layout + data set resident on one GPU.

    Parameters
    ----------
    mode : {'behavioral', 'behavioral_cov', 'meancentered'}
    S, B, T : int
        Rows, features, behaviours (``T`` ignored for mean-centred PLS)
    groups : list of int
        Subjects per group; rows are ordered group -> condition -> subject
        (pyls/structures.py:37-44)
    n_cond : int
    mean_centering : {0, 1, 2}
    device : int
        CUDA ordinal
    """

    def __init__(self, mode, S, B, T, groups, n_cond=1, mean_centering=0,
                 device=0, workspace_bytes=None, n_components=0,
                 gemm_backend='auto', gemm_slices=6):
        if not torch.cuda.is_available():
            raise RuntimeError('pypyls_b200 needs a CUDA device (B200, '
                               'sm_100a); there is no CPU fallback.')
        self._lib = _cabi.lib()
        self.device = torch.device('cuda', int(device))
        self.mode = mode
        groups = [int(g) for g in groups]
        self.S, self.B, self.T = int(S), int(B), int(T)
        self.groups, self.n_cond = groups, int(n_cond)
        self.J = len(groups) * self.n_cond
        self.K = self.J * self.T if mode != 'meancentered' else self.J
        if mode == 'regression':
            self.K = int(n_components)
        # compute.svd keeps min(K, B) latent variables (pyls/compute.py:36-50)
        self.L = self.K if mode == 'regression' else min(self.K, self.B)
        pool = _FREE_HANDLES.setdefault(self.device.index, [])
        if pool:
            self._h, idle = pool.pop()
            # the previous borrower's kernels may still be queued (on any
            # stream); plsb_configure uploads its tables synchronously and the
            # workspaces are shared, so wait for them here
            idle.synchronize()
        else:
            self._h = C.c_void_p(0)
            with torch.cuda.device(self.device):
                _cabi.check(self._lib.plsb_create(C.byref(self._h),
                                                  self.device.index))
        # pooled handles keep their settings: always (re)set them
        if not workspace_bytes:
            total = torch.cuda.get_device_properties(self.device).total_memory
            workspace_bytes = min(DEFAULT_WORKSPACE, total // 5)
        _cabi.check(self._lib.plsb_set_workspace_limit(
            self._h, int(workspace_bytes)))
        _cabi.check(self._lib.plsb_timing_enable(self._h, 0))
        self.set_gemm_backend(gemm_backend, gemm_slices)
        garr = (C.c_int * len(groups))(*groups)
        _cabi.check(self._lib.plsb_configure(
            self._h, MODES[mode], self.S, self.B, self.T, len(groups), garr,
            self.n_cond, int(mean_centering), int(n_components)))

    # -- plumbing ----------------------------------------------------------
    def set_gemm_backend(self, backend='auto', n_slices=6):
        """Kernel of the cross-covariance contraction: 'auto' = int8 slice
        GEMM on the tcgen05 tensor cores where it applies (``n_slices`` digit
        planes per operand: 6 -> ~1e-13, 7 -> ~1e-15 of |a||x| per entry),
        'dmma' = FP64 DMMA kernel everywhere."""
        code = {'auto': _cabi.PLSB_GEMM_AUTO, 'dmma': _cabi.PLSB_GEMM_DMMA}[backend]
        _cabi.check(self._lib.plsb_set_gemm_backend(self._h, code,
                                                    int(n_slices)))

    def gemm_work(self, reset=True):
        """(int8 multiply-accumulates of the slice GEMM, FP64 flop of the DMMA
        kernel) executed since the last reset."""
        a, b = C.c_double(0.0), C.c_double(0.0)
        _cabi.check(self._lib.plsb_gemm_work(self._h, C.byref(a), C.byref(b),
                                             int(bool(reset))))
        return float(a.value), float(b.value)

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            # the next borrower waits for everything queued so far on this
            # engine's stream before it touches the handle's buffers
            idle = torch.cuda.Event()
            idle.record(torch.cuda.current_stream(self.device))
            _FREE_HANDLES.setdefault(self.device.index, []).append(
                (self._h, idle))
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f64(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.device)

    def to_device(self, a, dtype=torch.float64):
        """Host array (NumPy or tensor) or device tensor -> contiguous device
        tensor of `dtype`."""
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype,
                        non_blocking=True).contiguous()
        a = np.ascontiguousarray(a)
        return torch.from_numpy(a).to(device=self.device, dtype=dtype,
                                      non_blocking=True).contiguous()

    def to_device_indices(self, samples):
        """Resampling table as the reference stores it, (S, n) integers
        (pyls/base.py:35,107), -> (n, S) int32 on the device."""
        if isinstance(samples, torch.Tensor) and samples.is_cuda:
            if samples.dtype != torch.int32 or samples.shape[-1] != self.S:
                raise ValueError('device index tables must be (n, S) int32')
            return samples.contiguous()
        samples = np.asarray(samples)
        if samples.ndim == 1:
            samples = samples[:, None]
        if samples.ndim != 2 or samples.shape[0] != self.S:
            raise ValueError('Resampling array must have shape ({}, n); got '
                             '{}'.format(self.S, samples.shape))
        if samples.size and (samples.min() < 0 or samples.max() >= self.S):
            raise ValueError('Resampling array has indices outside '
                             '[0, {})'.format(self.S))
        host = np.ascontiguousarray(samples.T.astype(np.int32))
        return torch.from_numpy(host).to(self.device)

    @property
    def launch_count(self):
        return int(self._lib.plsb_launch_count(self._h))

    # -- data / original decomposition ---------------------------------------
    def set_data(self, X, Y=None):
        X = self.to_device(X)
        if tuple(X.shape) != (self.S, self.B):
            raise ValueError('X must have shape ({}, {})'.format(self.S, self.B))
        if self.mode != 'meancentered':
            Y = self.to_device(Y)
            if tuple(Y.shape) != (self.S, self.T):
                raise ValueError('Y must have shape ({}, {})'
                                 .format(self.S, self.T))
        else:
            Y = None
        _cabi.check(self._lib.plsb_set_data(self._h, _ptr(X), _ptr(Y),
                                            self._stream()))
        return self

    def decompose(self):
        """Original decomposition -> (U (B,L), d (L,), V (K,L)) on the device;
        also installs it as the original for the resampling drivers."""
        U, d, V = self._f64(self.B, self.L), self._f64(self.L), \
            self._f64(self.K, self.L)
        _cabi.check(self._lib.plsb_decompose(self._h, _ptr(U), _ptr(d),
                                             _ptr(V), self._stream()))
        return U, d, V

    def set_original(self, U, d, V):
        U, d, V = self.to_device(U), self.to_device(d), self.to_device(V)
        if d.ndim == 2:
            d = torch.diagonal(d).contiguous()
        if tuple(U.shape) != (self.B, self.L) or \
                tuple(V.shape) != (self.K, self.L) or d.numel() != self.L:
            raise ValueError('original decomposition has the wrong shape')
        _cabi.check(self._lib.plsb_set_original(self._h, _ptr(U), _ptr(d),
                                                _ptr(V), self._stream()))
        return self

    def project_scores(self, U):
        U = self.to_device(U)
        out = self._f64(self.S, U.shape[1])
        _cabi.check(self._lib.plsb_project_scores(
            self._h, _ptr(U), int(U.shape[1]), _ptr(out), self._stream()))
        return out

    # -- index generation -----------------------------------------------------
    def _gen(self, fn, seed, first, count):
        idx = torch.empty((count, self.S), dtype=torch.int32,
                          device=self.device)
        n_ex = C.c_int(0)
        _cabi.check(fn(self._h, C.c_uint64(int(seed) & (2 ** 64 - 1)),
                       int(first), int(count), _ptr(idx), C.byref(n_ex),
                       self._stream()))
        return idx, int(n_ex.value)

    def gen_perm_indices(self, seed, count, first=0):
        """(count, S) int32 permutation table for resample ids
        [first, first+count) and the number of columns that hit the 500-draw
        cap (pyls/base.py:10-79)."""
        return self._gen(self._lib.plsb_gen_perm_indices, seed, first, count)

    def gen_boot_indices(self, seed, count, first=0):
        """Bootstrap table (pyls/base.py:82-159); see gen_perm_indices."""
        return self._gen(self._lib.plsb_gen_boot_indices, seed, first, count)

    def gen_split_masks(self, seed, count, n_split, first=0, train_fraction=0.5):
        """(count, n_split, S) int32 half / half masks for the sets (one per
        permutation) [first, first+count) and the number of masks that hit
        the 500-draw cap (gen_splits, pyls/base.py:162-229)."""
        masks = torch.empty((count, n_split, self.S), dtype=torch.int32,
                            device=self.device)
        n_ex = C.c_int(0)
        _cabi.check(self._lib.plsb_gen_split_masks(
            self._h, C.c_uint64(int(seed) & (2 ** 64 - 1)), int(first),
            int(count), int(n_split), float(train_fraction), _ptr(masks),
            C.byref(n_ex), self._stream()))
        return masks, int(n_ex.value)

    # -- resampling drivers ---------------------------------------------------
    def run_perms(self, idx, rotate=True):
        """Permuted singular values, (count, L) on the device
        (BasePLS.permutation, pyls/base.py:601-712)."""
        idx = self.to_device_indices(idx)
        out = self._f64(idx.shape[0], self.L)
        _cabi.check(self._lib.plsb_run_perms(
            self._h, _ptr(idx), int(idx.shape[0]), int(bool(rotate)),
            _ptr(out), self._stream()))
        return out

    def run_perms_gram(self, idx):
        """Rotated permuted singular values through the S x S Gram matrix of
        the data (sample-space fast path; same result as run_perms(rotate=True)
        without the cross-covariance GEMM)."""
        idx = self.to_device_indices(idx)
        out = self._f64(idx.shape[0], self.L)
        _cabi.check(self._lib.plsb_run_perms_gram(
            self._h, _ptr(idx), int(idx.shape[0]), _ptr(out), self._stream()))
        return out

    def run_perms_prepermuted(self, Yperm, rotate=True):
        """Permuted singular values (count, L) for pre-permuted behaviour
        matrices Yperm (count, S, T) -- `permsamples` with permindices=False
        (pyls/base.py:636-639, 689-692); X stays in place."""
        Yperm = self.to_device(Yperm)
        if Yperm.dim() != 3 or tuple(Yperm.shape[1:]) != (self.S, self.T):
            raise ValueError('pre-permuted Y must have shape (n, {}, {}); got '
                             '{}'.format(self.S, self.T, tuple(Yperm.shape)))
        out = self._f64(Yperm.shape[0], self.L)
        _cabi.check(self._lib.plsb_run_perms_prepermuted(
            self._h, _ptr(Yperm), int(Yperm.shape[0]), int(bool(rotate)),
            _ptr(out), self._stream()))
        return out

    def crossval(self, train):
        """Cross-validated Pearson r and R^2, (count, T) each, for train / test
        masks `train` (S, count) bool as gen_splits returns them
        (BehavioralPLS.crossval, pyls/types/behavioral.py:82-170)."""
        train = np.asarray(train)
        if train.ndim != 2 or train.shape[0] != self.S:
            raise ValueError('train / test masks must have shape ({}, n); got '
                             '{}'.format(self.S, train.shape))
        n_test = (train == 0).sum(axis=0)
        if n_test.min() < 2:
            raise ValueError('every split needs at least two held-out rows')
        mask = self.to_device(np.ascontiguousarray(train.T), dtype=torch.int32)
        n = int(mask.shape[0])
        r, r2 = self._f64(n, self.T), self._f64(n, self.T)
        _cabi.check(self._lib.plsb_crossval(
            self._h, _ptr(mask), n, int(n_test.max()), _ptr(r), _ptr(r2),
            self._stream()))
        return r, r2

    def split_half(self, masks, idx=None, Yperm=None, use_original=False):
        """Split-half correlations (ucorr, vcorr), (count, L) each on the
        device (BasePLS.split_half, pyls/base.py:714-770) of `count` data sets:
        the permutations `idx` ((S, count) host table or (count, S) device
        block), the pre-permuted behaviour matrices `Yperm` (count, S, T), or
        -- with neither -- the un-permuted data (count = 1).  `masks`:
        (count, n_split, S) half / half masks, non-zero = first half.
        ``use_original``: score against the installed original decomposition
        instead of every data set's own."""
        if isinstance(masks, torch.Tensor):
            masks = masks.to(device=self.device, dtype=torch.int32).contiguous()
        else:
            masks = self.to_device(np.ascontiguousarray(masks),
                                   dtype=torch.int32)
        if masks.dim() != 3 or masks.shape[2] != self.S:
            raise ValueError('split-half masks must have shape (count, '
                             'n_split, {}); got {}'.format(
                                 self.S, tuple(masks.shape)))
        n, n_split = int(masks.shape[0]), int(masks.shape[1])
        if idx is not None:
            idx = self.to_device_indices(idx)
            if int(idx.shape[0]) != n:
                raise ValueError('one set of masks per permutation is needed')
        if Yperm is not None:
            Yperm = self.to_device(Yperm)
            if tuple(Yperm.shape) != (n, self.S, self.T):
                raise ValueError('pre-permuted Y must have shape ({}, {}, {})'
                                 .format(n, self.S, self.T))
        if idx is None and Yperm is None and n != 1:
            raise ValueError('the un-permuted data is one data set')
        uc, vc = self._f64(n, self.L), self._f64(n, self.L)
        _cabi.check(self._lib.plsb_split_half(
            self._h, _ptr(idx), _ptr(Yperm), n, _ptr(masks), n_split,
            int(bool(use_original)), _ptr(uc), _ptr(vc), self._stream()))
        return uc, vc

    def boot_distrib(self, idx):
        """Bootstrap distribution (count, K, L) alone (gen_distrib needs only
        the original weights): ready, and on its way to the host, before the
        cross-covariance work of run_boots(..., want_distrib=False) starts."""
        idx = self.to_device_indices(idx)
        n = int(idx.shape[0])
        distrib = self._f64(n, self.K, self.L)
        _cabi.check(self._lib.plsb_boot_distrib(
            self._h, _ptr(idx), n, _ptr(distrib), self._stream()))
        return distrib

    def run_boots(self, idx, u_sum=None, u_square=None, want_distrib=True):
        """(distrib (count,K,L), u_sum (B,L), u_square (B,L)) on the device
        (BasePLS.bootstrap, pyls/base.py:439-576); ``want_distrib=False``
        leaves the distribution out (None)."""
        idx = self.to_device_indices(idx)
        n = int(idx.shape[0])
        distrib = self._f64(n, self.K, self.L) if want_distrib else None
        if u_sum is None:
            u_sum = torch.zeros((self.B, self.L), dtype=torch.float64,
                                device=self.device)
            u_square = torch.zeros_like(u_sum)
        _cabi.check(self._lib.plsb_run_boots(
            self._h, _ptr(idx), n, _ptr(distrib), _ptr(u_sum), _ptr(u_square),
            self._stream()))
        return distrib, u_sum, u_square

    def run_boots_streamed(self, idx, wait=True, host=None):
        """run_boots in blocks of the library's internal pass size; the block's
        slice of `distrib` is copied to pinned host memory on a side stream
        while the next block computes.  Returns (distrib on the device,
        the same (count, K, L) on the host, u_sum, u_square); with
        ``wait=False`` the main stream is not made to wait for the copies and
        the event that marks the last one is returned as a fifth item.  `host`:
        optional pinned (count, K, L) tensor (e.g. this rank's slice of a
        larger buffer) to fill instead of a new one."""
        idx = self.to_device_indices(idx)
        n = int(idx.shape[0])
        distrib = self._f64(n, self.K, self.L)
        if host is None:
            host = torch.empty((n, self.K, self.L), dtype=torch.float64,
                               pin_memory=True)
        u_sum = torch.zeros((self.B, self.L), dtype=torch.float64,
                            device=self.device)
        u_square = torch.zeros_like(u_sum)
        chunk = max(1, int(self._lib.plsb_boot_chunk(self._h, n))) if n else 1
        main = torch.cuda.current_stream(self.device)
        side = copy_stream(self.device)
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            _cabi.check(self._lib.plsb_run_boots(
                self._h, _ptr(idx[a:b]), b - a, _ptr(distrib[a:b]),
                _ptr(u_sum), _ptr(u_square), self._stream()))
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(side):
                side.wait_event(done)
                host[a:b].copy_(distrib[a:b], non_blocking=True)
        if wait:
            main.wait_stream(side)
            return distrib, host, u_sum, u_square
        last = torch.cuda.Event()
        last.record(side)
        return distrib, host, u_sum, u_square, last

    def crosscov(self, idx=None, bootstrap=False):
        """Cross-covariance matrices (count, K, B) of the given resamples;
        ``idx=None`` gives the un-resampled matrix."""
        if idx is None:
            n, p = 1, C.c_void_p(0)
        else:
            idx = self.to_device_indices(idx)
            n, p = int(idx.shape[0]), _ptr(idx)
        out = self._f64(n, self.K, self.B)
        _cabi.check(self._lib.plsb_crosscov(self._h, p, n, int(bool(bootstrap)),
                                            _ptr(out), self._stream()))
        return out

    # -- SIMPLS (pls_regression) -------------------------------------------------
    def _omega(self, omega, n):
        if omega is None:
            if self.T > 11:
                raise ValueError('SIMPLS with more than 11 behaviours needs '
                                 'the Gaussian test matrices')
            return None
        omega = self.to_device(omega)
        if tuple(omega.shape) != (n, self.T, 11):
            raise ValueError('omega must have shape ({}, {}, 11)'
                             .format(n, self.T))
        return omega

    def gen_gaussian_tables(self, first, count):
        """(count, T, 11) Gaussian test matrices of resamples
        [first, first + count) on the device (NumPy's legacy
        ``RandomState(i).normal(size=(T, 11))`` replayed per resample)."""
        out = self._f64(count, self.T, 11)
        _cabi.check(self._lib.plsb_gen_gaussian_tables(
            self._h, int(first), int(count), self.T, _ptr(out),
            self._stream()))
        return out

    def simpls_decompose(self, omega=None):
        """Original SIMPLS decomposition -> (x_weights (B,L), pctvar_y (L,)).
        `omega` (L, T, 11): one Gaussian test matrix per component."""
        omega = self._omega(omega, self.L)
        xw, pct = self._f64(self.B, self.L), self._f64(self.L)
        _cabi.check(self._lib.plsb_simpls_decompose(
            self._h, _ptr(omega), _ptr(xw), _ptr(pct), self._stream()))
        return xw, pct

    def simpls_set_row_mask(self, valid_x, valid_y):
        """Marks the rows of X and of Y (False / 0, shape (S,) each) that are
        missing altogether -- what the reference's get_mask drops
        (pyls/types/regression.py:48-53); call after set_data (with those rows
        zero-filled), before simpls_decompose."""
        dev = []
        for v in (valid_x, valid_y):
            v = np.asarray(v).astype(np.int32).ravel()
            if v.size != self.S:
                raise ValueError('row masks must have {} entries'
                                 .format(self.S))
            dev.append(torch.from_numpy(np.ascontiguousarray(v))
                       .to(self.device))
        _cabi.check(self._lib.plsb_simpls_set_row_mask(
            self._h, _ptr(dev[0]), _ptr(dev[1]), self._stream()))
        return self

    def simpls_set_original(self, x_weights):
        xw = self.to_device(x_weights)
        _cabi.check(self._lib.plsb_simpls_set_original(self._h, _ptr(xw),
                                                       self._stream()))
        return self

    def simpls_run_perms(self, idx, omega=None):
        """pctvar in Y of every permutation, (count, L) on the device."""
        idx = self.to_device_indices(idx)
        n = int(idx.shape[0])
        omega = self._omega(omega, n)
        out = self._f64(n, self.L)
        _cabi.check(self._lib.plsb_simpls_run_perms(
            self._h, _ptr(idx), n, _ptr(omega), _ptr(out), self._stream()))
        return out

    def simpls_run_boots(self, idx, omega=None, yres=None, u_sum=None,
                         u_square=None):
        """(distrib (count,T,L), u_sum (B,L), u_square (B,L), pctvar
        (count,L)) of the bootstraps, on the device.  `yres` (count, S, T):
        a behaviour matrix of its own for every bootstrap (three-dimensional Y
        aggregated per bootstrap, pyls/types/regression.py:308-310)."""
        idx = self.to_device_indices(idx)
        n = int(idx.shape[0])
        omega = self._omega(omega, n)
        pct = self._f64(n, self.L)
        distrib = self._f64(n, self.T, self.L)
        if u_sum is None:
            u_sum = torch.zeros((self.B, self.L), dtype=torch.float64,
                                device=self.device)
            u_square = torch.zeros_like(u_sum)
        if yres is None:
            _cabi.check(self._lib.plsb_simpls_run_boots(
                self._h, _ptr(idx), n, _ptr(omega), _ptr(pct), _ptr(distrib),
                _ptr(u_sum), _ptr(u_square), self._stream()))
        else:
            yres = self.to_device(yres)
            if tuple(yres.shape) != (n, self.S, self.T):
                raise ValueError('per-bootstrap Y must have shape ({}, {}, {})'
                                 .format(n, self.S, self.T))
            _cabi.check(self._lib.plsb_simpls_run_boots_yres(
                self._h, _ptr(idx), n, _ptr(omega), _ptr(yres), _ptr(pct),
                _ptr(distrib), _ptr(u_sum), _ptr(u_square), self._stream()))
        return distrib, u_sum, u_square, pct

    # -- statistics -----------------------------------------------------------
    def perm_pvals(self, d_perm, d_orig):
        d_perm, d_orig = self.to_device(d_perm), self.to_device(d_orig)
        out = self._f64(d_perm.shape[1])
        _cabi.check(self._lib.plsb_perm_pvals(
            self._h, _ptr(d_perm), int(d_perm.shape[0]), int(d_perm.shape[1]),
            _ptr(d_orig), _ptr(out), self._stream()))
        return out

    def percentile(self, distrib, q_lo, q_hi):
        """numpy-'linear' percentiles over axis 0 of a (count, ...) tensor."""
        distrib = self.to_device(distrib)
        n = int(distrib.shape[0])
        series = int(distrib.numel() // max(n, 1))
        lo, hi = self._f64(*distrib.shape[1:]), self._f64(*distrib.shape[1:])
        _cabi.check(self._lib.plsb_percentile(
            self._h, _ptr(distrib), n, series, float(q_lo), float(q_hi),
            _ptr(lo), _ptr(hi), self._stream()))
        return lo, hi

    def percentile_series(self, series, q_lo, q_hi):
        """The same percentiles for a series-major (n_series, count) tensor
        (rows may be padded: stride(0) >= count)."""
        if series.dim() != 2 or series.stride(1) != 1:
            raise ValueError('series-major input must be 2-D with unit column '
                             'stride')
        n_series, n = int(series.shape[0]), int(series.shape[1])
        lo, hi = self._f64(n_series), self._f64(n_series)
        _cabi.check(self._lib.plsb_percentile_series(
            self._h, _ptr(series), n_series, n, int(series.stride(0)),
            float(q_lo), float(q_hi), _ptr(lo), _ptr(hi), self._stream()))
        return lo, hi

    def transpose(self, a):
        """(rows, cols) -> contiguous (cols, rows) on the device."""
        a = self.to_device(a)
        rows, cols = int(a.shape[0]), int(a.numel() // max(int(a.shape[0]), 1))
        out = self._f64(cols, rows)
        _cabi.check(self._lib.plsb_transpose(self._h, _ptr(a), rows, cols,
                                             _ptr(out), self._stream()))
        return out

    def boot_ratio(self, bs, u_sum, u_square, n_boot, add_orig):
        bs = self.to_device(bs)
        bsr, se = torch.empty_like(bs), torch.empty_like(bs)
        _cabi.check(self._lib.plsb_boot_ratio(
            self._h, _ptr(bs), _ptr(u_sum), _ptr(u_square), int(bs.numel()),
            int(n_boot), int(bool(add_orig)), _ptr(bsr), _ptr(se),
            self._stream()))
        return bsr, se

    # -- primitives (unit tests) ----------------------------------------------
    def dgemm(self, A, X):
        A, X = self.to_device(A), self.to_device(X)
        M, Kd = A.shape
        N = X.shape[1]
        out = self._f64(M, N)
        _cabi.check(self._lib.plsb_dgemm(self._h, _ptr(A), _ptr(X), int(M),
                                         int(N), int(Kd), _ptr(out),
                                         self._stream()))
        return out

    def gemm_probe(self, variant, M, N, Kd, k_valid=0, scale_div=10, iters=3):
        """Mean milliseconds per launch of a cross-covariance GEMM variant on
        synthetic operands (0 store, 1 scaled store, 2 row sums of squares)."""
        ms = C.c_double(0.0)
        _cabi.check(self._lib.plsb_gemm_probe(
            self._h, int(variant), int(M), int(N), int(Kd), int(k_valid),
            int(scale_div), int(iters), C.byref(ms), self._stream()))
        return float(ms.value)

    def small_decomp(self, G, H, d_orig=None):
        G, H = self.to_device(G), self.to_device(H)
        d_orig = None if d_orig is None else self.to_device(d_orig)
        n, K = int(G.shape[0]), int(G.shape[1])
        L = int(H.shape[2])
        M, lam = self._f64(n, K, L), self._f64(n, K)
        _cabi.check(self._lib.plsb_small_decomp(
            self._h, _ptr(G), _ptr(H), n, K, L, _ptr(d_orig), _ptr(M),
            _ptr(lam), self._stream()))
        return M, lam

    def gram_proj(self, R, Uo=None):
        """G[r] = R[r] R[r]^T and H[r] = R[r] Uo for a stack R (n, K, B)."""
        R = self.to_device(R)
        n, K, B = (int(x) for x in R.shape)
        Uo = None if Uo is None else self.to_device(Uo)
        L = 0 if Uo is None else int(Uo.shape[1])
        G = self._f64(n, K, K)
        H = None if Uo is None else self._f64(n, K, L)
        _cabi.check(self._lib.plsb_gram_proj(
            self._h, _ptr(R), n, K, B, _ptr(Uo), L, _ptr(G), _ptr(H),
            self._stream()))
        return G, H

    def accum_u(self, R, M):
        """sum_r R[r]^T M[r] and sum_r (R[r]^T M[r])^2 for stacks R (n, K, B),
        M (n, K, L)."""
        R, M = self.to_device(R), self.to_device(M)
        n, K, B = (int(x) for x in R.shape)
        L = int(M.shape[2])
        us = torch.zeros((B, L), dtype=torch.float64, device=self.device)
        uq = torch.zeros_like(us)
        _cabi.check(self._lib.plsb_accum_u(
            self._h, _ptr(R), n, K, B, _ptr(M), L, _ptr(us), _ptr(uq),
            self._stream()))
        return us, uq

    # -- per-kernel-class timing (bench) ---------------------------------------
    def timing_enable(self, on=True):
        _cabi.check(self._lib.plsb_timing_enable(self._h, int(bool(on))))

    def timing_read(self):
        """{class name: (milliseconds, launches)} since the previous read."""
        n = int(self._lib.plsb_timing_classes())
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        _cabi.check(self._lib.plsb_timing_read(self._h, ms, cnt))
        return {self._lib.plsb_timing_class_name(i).decode(): (ms[i], cnt[i])
                for i in range(n)}
