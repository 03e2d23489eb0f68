"""cProfile of one end-to-end front-end call at BASELINE config 2 (run on the GPU box)."""
import cProfile, pstats, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pypyls_b200 as pyls

rs = np.random.RandomState(1234)
X, Y = rs.rand(80, 10000), rs.rand(80, 10)
Xh, Yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(Y).pin_memory()
kw = dict(groups=[20, 20], n_cond=2, n_perm=5000, n_boot=5000, verbose=False)
for i in range(2):
    pyls.behavioral_pls(Xh, Yh, seed=i, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
pyls.behavioral_pls(Xh, Yh, seed=7, **kw)
torch.cuda.synchronize()
pr.disable()
print('one call: %.1f ms' % (1e3 * (time.perf_counter() - t0)))
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
