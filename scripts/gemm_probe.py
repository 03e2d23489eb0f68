"""Times the variants of the cross-covariance GEMM in isolation (plsb_gemm_probe) under
several tuning-knob settings.  Usage: python scripts/gemm_probe.py [M] [N] [Kd] [k_valid]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypyls_b200.engine import ResamplingEngine  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
Kd = int(sys.argv[3]) if len(sys.argv) > 3 else 208
kv = int(sys.argv[4]) if len(sys.argv) > 4 else 200
eng = ResamplingEngine('behavioral', 16, 64, 2, [16], 1, device=0)
names = {0: 'store', 1: 'store+scale', 2: 'rowsumsq'}
settings = [s for s in os.environ.get('PROBE_SETTINGS', '').split(';') if s] or [
    'PLSB_GEMM_V=1', 'PLSB_GEMM_V=2',
    'PLSB_GEMM_V=1 PLSB_GEMM_SMALL_TILE=1 PLSB_GEMM_ROWSQ_SMALL=1',
    'PLSB_GEMM_V=2 PLSB_GEMM_SMALL_TILE=1 PLSB_GEMM_ROWSQ_SMALL=1']
flop = 2.0 * M * N * kv
for setting in settings:
    keys = []
    for kvp in setting.split():
        k, v = kvp.split('=', 1)
        os.environ[k] = v
        keys.append(k)
    out = []
    for variant in (0, 1, 2):
        ms = eng.gemm_probe(variant, M, N, Kd, k_valid=kv, iters=3)
        out.append('%s %.2f ms %.1f TF' % (names[variant], ms, flop / ms / 1e9))
    print('%-70s | %s' % (setting, ' | '.join(out)), flush=True)
    for k in keys:
        del os.environ[k]
