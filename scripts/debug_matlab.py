import sys, os, warnings
import numpy as np
warnings.filterwarnings('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_golden
from oracle import pls_oracle as po
import pypyls_b200 as pyls
from pypyls_b200.engine import ResamplingEngine

ins, ref = load_golden('matlab_bpls_onegroup_onecond_nosplit')
X, Y = ins['X'], ins['Y']
spec = po._Spec('behavioral', ins['groups'], ins['n_cond'])
R0 = po.gen_covcorr(spec, X, Y)
U, d, Vt = np.linalg.svd(R0.T, full_matrices=False); V = Vt.T
eng = ResamplingEngine('behavioral', 40, 75, 25, [40], 1)
eng.set_data(X, Y)
Rg = eng.crosscov().cpu().numpy()[0]
print('R0 err', np.abs(Rg - R0).max())
Ug, dg, Vg = [t.cpu().numpy() for t in eng.decompose()]
sg = np.sign((Vg * V).sum(0))
print('d err', np.abs(dg - d).max(), 'V err', np.abs(Vg * sg - V).max(), 'U err', np.abs(Ug * sg - U).max())
print('V ortho', np.abs(Vg.T @ Vg - np.eye(25)).max(), 'col norms', np.abs(np.linalg.norm(Vg, axis=0) - 1).max())
ps = ins['permsamples']
dp = eng.run_perms(ps[:, :8], rotate=True).cpu().numpy()
for i in range(3):
    Rp = po.gen_covcorr(spec, X, Y[ps[:, i]])
    want = np.linalg.norm(Rp.T @ Vg, axis=0)
    print(i, 'perm err vs identity(own V)', np.abs(dp[i] - want).max(), 'vs golden', np.abs(dp[i] - ref['py_perm_singval'][:, i]).max())
Rpg = eng.crosscov(ps[:, :3]).cpu().numpy()
for i in range(3):
    print(i, 'Rperm err', np.abs(Rpg[i] - po.gen_covcorr(spec, X, Y[ps[:, i]])).max())
dn = eng.run_perms(ps[:, :3], rotate=False).cpu().numpy()
for i in range(3):
    Rp = po.gen_covcorr(spec, X, Y[ps[:, i]])
    print(i, 'norot err', np.abs(dn[i] - np.linalg.svd(Rp, compute_uv=False)).max())
