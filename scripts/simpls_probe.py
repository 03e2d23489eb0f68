"""Runs a small batch of cfg4-shaped SIMPLS resamples (for ncu)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pypyls_b200.engine import ResamplingEngine
from pypyls_b200.types.regression import gaussian_tables
rs = np.random.RandomState(1234)
S, B, T, L, n = 500, 5000, 20, 10, 296
X, Y = rs.rand(S, B), rs.rand(S, T)
eng = ResamplingEngine('regression', S, B, T, [S], 1, n_components=L)
eng.set_data(X - X.mean(0), Y - Y.mean(0))
om0 = np.stack([rs.normal(size=(T, 11)) for _ in range(L)])
eng.simpls_decompose(om0)
om = eng.to_device(gaussian_tables(range(n), T))
idx, _ = eng.gen_boot_indices(1, n)
for _ in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); eng.simpls_run_boots(idx, om); e1.record(); torch.cuda.synchronize()
    print('boots ms', e0.elapsed_time(e1))
