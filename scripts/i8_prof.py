"""Wait-cycle profile of the int8 slice GEMM (PLSB_I8_PROF=1): python scripts/i8_prof.py [M] [N] [slices,..]
(build with PLSB_NVCC_EXTRA=-DPLSB_I8_EPI_PROF for the phases of the epilogue as well)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['PLSB_I8_PROF'] = '1'
from pypyls_b200.engine import ResamplingEngine  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
slices = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [6, 5, 7]
eng = ResamplingEngine('behavioral', 16, 64, 2, [16], 1, device=0)
for s in slices:
    eng.set_gemm_backend('auto', s)
    for variant in (2, 1, 0):
        if variant < 2 and M > 30000:
            continue
        ms = eng.gemm_probe(variant, M, N, 208, k_valid=200, iters=1)
        print('slices %d variant %d: %.2f ms' % (s, variant, ms), flush=True)
