"""GPU timeline of one end-to-end front-end call: busy time vs span and the largest idle
gaps (torch.profiler / CUPTI; run on the GPU box).
Usage: python scripts/e2e_timeline.py [cfg2|cfg5] [n_each]"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import pypyls_b200 as pyls

import time
rs = np.random.RandomState(1234)
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
if cfg == 'cfg5':
    X, Y = rs.rand(200, 100000), rs.rand(200, 10)
    n_each = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    kw = dict(n_perm=n_each, n_boot=n_each, verbose=False)
else:
    X, Y = rs.rand(80, 10000), rs.rand(80, 10)
    n_each = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    kw = dict(groups=[20, 20], n_cond=2, n_perm=n_each, n_boot=n_each, verbose=False)
Xh, Yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(Y).pin_memory()
for i in range(3):
    pyls.behavioral_pls(Xh, Yh, seed=i, **kw)
torch.cuda.synchronize()
t_wall = time.perf_counter()
pyls.behavioral_pls(Xh, Yh, seed=5, **kw)
torch.cuda.synchronize()
print('wall clock of one call without the profiler: %.2f ms' % (1e3 * (time.perf_counter() - t_wall)))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pyls.behavioral_pls(Xh, Yh, seed=7, **kw)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy, cur_end, gaps = 0.0, t0, []
for e in ev:
    s, en = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, prev.name[:60], e.name[:60], (cur_end - t0) / 1e3))
    busy += max(0.0, en - max(s, cur_end))
    if en > cur_end:
        cur_end, prev = en, e
print('span %.2f ms, busy %.2f ms, idle %.2f ms, %d device events' %
      ((t1 - t0) / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3, len(ev)))
agg = {}
for e in ev:
    agg[e.name[:70]] = agg.get(e.name[:70], 0.0) + (e.time_range.end - e.time_range.start)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:14]:
    print('%9.1f us  %s' % (v, k))
for g in sorted(gaps, reverse=True)[:14]:
    print('gap %7.1f us at %6.2f ms  after %-60s before %s' % (g[0], g[3], g[1], g[2]))
