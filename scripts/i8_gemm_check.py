"""Correctness and speed of the int8 slice GEMM (gemm_i8.cu) against NumPy and the DMMA kernel.
Usage: python scripts/i8_gemm_check.py [quick]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pypyls_b200.engine import ResamplingEngine  # noqa: E402

eng = ResamplingEngine('behavioral', 16, 64, 2, [16], 1, device=0)
rng = np.random.RandomState(0)


def check(M, N, Kd, slices, scale=1.0):
    A = rng.randn(M, Kd) * scale
    X = rng.randn(Kd, N)
    ref = A @ X
    norm = np.sqrt((A ** 2).sum(1))[:, None] * np.sqrt((X ** 2).sum(0))[None, :]
    out = {}
    for backend in ('dmma', 'auto'):
        eng.set_gemm_backend(backend, slices)
        torch.cuda.synchronize()
        C = eng.dgemm(A, X).cpu().numpy()
        out[backend] = C
    e_d = np.abs(out['dmma'] - ref).max() / np.abs(ref).max()
    e_i = np.abs(out['auto'] - ref).max() / np.abs(ref).max()
    e_n = (np.abs(out['auto'] - ref) / norm).max()
    print('M=%d N=%d Kd=%d slices=%d: max err / max|C|  dmma %.2e  i8 %.2e   i8 normwise %.2e'
          % (M, N, Kd, slices, e_d, e_i, e_n), flush=True)
    return e_i


for slices in (6, 7, 5):
    check(128, 128, 200, slices)
    check(300, 1000, 200, slices)
    check(1000, 5000, 37, slices, scale=1e-3)
    check(4000, 20000, 224, slices)
# non-finite input propagates like a product would
eng.set_gemm_backend('auto', 6)
A = rng.randn(256, 100)
X = rng.randn(100, 256)
A[3, 5] = np.nan
X[7, 9] = np.nan
C = eng.dgemm(A, X).cpu().numpy()
ok = np.isnan(C[3]).all() and np.isnan(C[:, 9]).all() and np.isfinite(np.delete(np.delete(C, 3, 0), 9, 1)).all()
print('NaN propagation:', 'ok' if ok else 'WRONG', flush=True)

if len(sys.argv) > 1 and sys.argv[1] == 'quick':
    sys.exit(0)
names = {0: 'store', 1: 'store+scale', 2: 'rowsumsq'}
for (M, N) in ((25000, 100000), (100000, 100000)):
    flop = 2.0 * M * N * 200
    for backend, slices in (('dmma', 6), ('auto', 6), ('auto', 7), ('auto', 5)):
        eng.set_gemm_backend(backend, slices)
        out = []
        for variant in (0, 1, 2):
            if variant < 2 and M > 30000:
                continue
            ms = eng.gemm_probe(variant, M, N, 208, k_valid=200, iters=3)
            out.append('%s %.2f ms %.1f TF' % (names[variant], ms, flop / ms / 1e9))
        print('M=%d N=%d %s/%d | %s' % (M, N, backend, slices, ' | '.join(out)), flush=True)
