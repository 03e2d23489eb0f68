"""One chunk of permutations + bootstraps of a bench workload, for ncu captures:

    ncu --set full --import-source on --clock-control none \
        -k regex:'xcov_gemm|gram_proj|accum_u|eigen|rotation' -c 16 \
        -o gpurun_out/boots python scripts/profile_boots.py [cfg2] [n]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import WORKLOADS, make_data  # noqa: E402
from pypyls_b200.engine import ResamplingEngine  # noqa: E402

wname = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
w = WORKLOADS[wname]
X, Y = make_data(w)
eng = ResamplingEngine(w['kind'], w['S'], w['B'], w['T'], w['groups'], w['n_cond'], device=0)
eng.set_data(torch.from_numpy(X), torch.from_numpy(Y) if w['kind'] == 'behavioral' else None)
U, d, V = eng.decompose()
idx_p, _ = eng.gen_perm_indices(1234, n)
idx_b, _ = eng.gen_boot_indices(1234, n)
eng.run_perms(idx_p, rotate=True)
eng.run_boots(idx_b)
torch.cuda.synchronize()
print('done', wname, n)
