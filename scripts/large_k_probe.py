"""Per-kernel-class times of a large-K analysis (reference integration shapes)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from pypyls_b200.engine import ResamplingEngine
from oracle import pls_oracle as po

groups, n_cond = ([25, 25], 2) if len(sys.argv) < 2 else (eval(sys.argv[1]), int(sys.argv[2]))
rs = np.random.RandomState(1234)
X, Y = rs.rand(100, 1000), rs.rand(100, 100)
eng = ResamplingEngine('behavioral', 100, 1000, 100, groups, n_cond).set_data(X, Y)
eng.timing_enable(True)
def lap(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    print('%-22s %.3f s' % (name, time.perf_counter() - t0), {k: round(v[0], 1) for k, v in eng.timing_read().items() if v[1]}, flush=True)
    return out
lap('decompose', eng.decompose)
ps = po.gen_permsamp(groups, n_cond, 20, seed=1)
bs = po.gen_bootsamp(groups, n_cond, 10, seed=2)
lap('perms rotate', lambda: eng.run_perms(ps, rotate=True))
lap('perms no rotate', lambda: eng.run_perms(ps, rotate=False))
lap('boots', lambda: eng.run_boots(bs))
