# -*- coding: utf-8 -*-
"""
NumPy restatement of the int8 slice GEMM of pypyls_b200/csrc/gemm_i8.cu -- TEST
INFRASTRUCTURE ONLY (imported by tests/, never by the product).

The product evaluates ``compute.xcorr``'s ``Yn.T @ Xn`` (pyls/compute.py:92) for
thousands of stacked resamples as exact int8 digit-plane products on the tcgen05
tensor cores.  This file restates that arithmetic step by step with integer NumPy
arrays, in the order the kernel performs it, so that the CUDA result can be
compared BIT FOR BIT (integer work is exact, the FP64 recombination has a fixed
order) and the error against the FP64 product can be bounded without a GPU.
"""

import numpy as np


def digit_planes(V, S, axis):
    """Balanced base-256 digit planes of the rows (axis=1) / columns (axis=0) of V.

    Every row (column) gets the power of two 2**e > max|v| and is cut into S signed
    digits of q = rint(v * 2**(8S - 2 - e)):  q = sum_i dig[i] * 256**(S-1-i), every
    digit in [-128, 127] (quant_rows_kernel / quant_cols_kernel, digits_into).
    Returns (planes (S, *V.shape) int64, exponents e per row / column, nonzero mask).
    """
    Q = 8 * S - 2
    mx = np.max(np.abs(V), axis=axis, keepdims=True)
    live = mx > 0
    e = np.where(live, np.frexp(np.where(live, mx, 1.0))[1], 0)       # mx < 2**e
    q = np.rint(np.ldexp(V, (Q - e).astype(np.int64) * np.ones(V.shape, np.int64))).astype(np.int64)
    q = np.where(live, q, 0)
    planes = np.zeros((S,) + V.shape, np.int64)
    for i in range(S - 1, 0, -1):
        low = ((q + 128) & 0xFF) - 128                                # signed low byte
        planes[i] = low
        q = (q - low) >> 8
    planes[0] = q
    assert np.abs(planes).max() <= 128
    return planes, np.squeeze(e, axis=axis), np.squeeze(live, axis=axis)


def slice_gemm(A, X, S=6):
    """C = A @ X the way xcov_gemm_i8_kernel (STORE epilogue) computes it."""
    A, X = np.asarray(A, np.float64), np.asarray(X, np.float64)
    Q = 8 * S - 2
    pa, ea, la = digit_planes(A, S, axis=1)            # planes of the rows of A
    px, ex, lx = digit_planes(X, S, axis=0)            # planes of the columns of X
    # exact int32 plane products, pairs with i + j = d in one accumulator
    T = np.zeros((A.shape[0], X.shape[1]), np.float64)
    for d in range(S - 1, -1, -1):                     # the kernel drains d = S-1 first
        P = np.zeros(T.shape, np.int64)
        for i in range(d + 1):
            P += pa[i] @ px[d - i]
        assert np.abs(P).max() < 2 ** 31
        T = T + np.ldexp(P.astype(np.float64), 8 * (S - 1 - d))
    rs = np.where(la, np.ldexp(1.0, ea - Q), 0.0)                      # rscale
    cs = np.where(lx, np.ldexp(1.0, ex - Q + 8 * (S - 1)), 0.0)        # cscale
    return T * (rs[:, None] * cs[None, :])
