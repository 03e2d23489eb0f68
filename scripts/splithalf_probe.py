"""Times the split-half path (plsb_split_half) at BASELINE config-2 shape:
P permutations x n_split masks x 2 halves, masks and tables generated on the device."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pypyls_b200.engine import ResamplingEngine

P, n_split = int(sys.argv[1]) if len(sys.argv) > 1 else 200, \
    int(sys.argv[2]) if len(sys.argv) > 2 else 50
rs = np.random.RandomState(1234)
S, B, T, groups, n_cond = 80, 10000, 10, [20, 20], 2
X, Y = rs.rand(S, B), rs.rand(S, T)
eng = ResamplingEngine('behavioral', S, B, T, groups, n_cond)
eng.set_data(X, Y)
eng.decompose()
idx, _ = eng.gen_perm_indices(1, P)
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(True) for _ in range(3))
    e0.record()
    masks, _ = eng.gen_split_masks(7, P, n_split)
    e1.record()
    uc, vc = eng.split_half(masks, idx=idx)
    e2.record()
    torch.cuda.synchronize()
    ms = e1.elapsed_time(e2)
    print('masks %.2f ms; split_half %.1f ms: %d perms x %d masks -> %.0f halves/s, '
          '%.1f permutations/s' % (e0.elapsed_time(e1), ms, P, n_split,
                                   2e3 * P * n_split / ms, 1e3 * P / ms))
