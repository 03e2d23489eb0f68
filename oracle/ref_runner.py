# -*- coding: utf-8 -*-
"""
Runs the UNMODIFIED reference package (oracle/_ref, see oracle/build_ref.py)
through its own public API and stock code path -- BASELINE / TEST
INFRASTRUCTURE ONLY (bench.py's ``cpu_baseline`` leg and ``--impl reference``).

    python oracle/ref_runner.py --workload cfg5 --n-each 200 --n-proc 32 \
        --blas-threads 1 [--steps K --warmup W] [--target-s 6] [--save out.npz]

One "step" is one call of ``pyls.behavioral_pls`` / ``meancentered_pls`` /
``pls_regression`` with ``n_perm = n_boot = n_each`` on the synthetic workload
of bench.py (``RandomState(1234).rand``), ``n_proc`` joblib workers
(pyls/utils.py:252-279) and the BLAS thread count given (the reference's CI
pins it to 1, .travis.yml:22-23).  The call is the reference's stock path;
the shims are the ones SURVEY.md 8(c) lists and none touches arithmetic:
a stand-in ``h5py`` module when h5py is absent, ``permindices=True`` (HEAD
mis-handles the missing kwarg, pyls/base.py:628-639) and, for
``pls_regression`` only, the ``_single_perm`` signature adapter
(pyls/types/regression.py:329 vs pyls/base.py:646-648).  Wall-clock timers are
wrapped around ``gen_permsamp`` / ``gen_bootsamp`` / ``BasePLS.permutation`` /
``BasePLS.bootstrap`` so that index generation and the two resampling loops
are reported separately.

Prints ONE JSON line; ``--save`` stores the last step's outputs (tables,
permuted singular values, p-values, bootstrap distribution, CIs, bootstrap
ratios) for the element-wise parity block of bench.py.
"""

import argparse
import json
import os
import sys
import time


def _parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='cfg5')
    ap.add_argument('--n-each', type=int, default=0,
                    help='permutations and bootstraps per step (0: calibrate '
                         'from --target-s)')
    ap.add_argument('--target-s', type=float, default=6.0)
    ap.add_argument('--max-each', type=int, default=2000)
    ap.add_argument('--n-proc', type=int, default=1)
    ap.add_argument('--blas-threads', type=int, default=1,
                    help='0 = library default')
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--warmup', type=int, default=0)
    ap.add_argument('--seed', type=int, default=1234)
    ap.add_argument('--save', default=None)
    return ap.parse_args()


ARGS = _parse() if __name__ == '__main__' else None
if ARGS is not None and ARGS.blas_threads > 0:
    for var in ('OPENBLAS_NUM_THREADS', 'OMP_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[var] = str(ARGS.blas_threads)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import warnings  # noqa: E402

import numpy as np  # noqa: E402

TIMERS = {}


def _timed(name, fn):
    def wrapper(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            TIMERS[name] = TIMERS.get(name, 0.0) + time.perf_counter() - t0
    return wrapper


def load_reference():
    """Imports the reference package from oracle/_ref and applies the shims."""
    from oracle import build_ref
    if not build_ref.available():
        raise ImportError('oracle/_ref is missing: run oracle/build_ref.py '
                          'where /root/reference exists')
    for p in reversed(build_ref.paths()):
        sys.path.insert(0, p)
    warnings.filterwarnings('ignore')
    import pyls
    from pyls import base
    from pyls.types.regression import PLSRegression
    if not getattr(PLSRegression, '_b200_adapter', False):
        orig = PLSRegression._single_perm

        def _single_perm(self, X, Y, samples, use_permind=True, groups=None,
                         original=None, seed=None):
            return orig(self, X, Y, inds=samples, groups=groups,
                        original=original, seed=seed)
        PLSRegression._single_perm = _single_perm
        PLSRegression._b200_adapter = True
        base.gen_permsamp = _timed('gen_permsamp', base.gen_permsamp)
        base.gen_bootsamp = _timed('gen_bootsamp', base.gen_bootsamp)
        base.BasePLS.permutation = _timed('permutation',
                                          base.BasePLS.permutation)
        base.BasePLS.bootstrap = _timed('bootstrap', base.BasePLS.bootstrap)
    return pyls


def call(pyls, w, X, Y, n_each, n_proc, seed, **extra):
    kw = dict(n_perm=n_each, n_boot=n_each, seed=seed, verbose=False,
              n_proc=n_proc, permindices=True)
    kw.update(extra)
    if w['kind'] == 'regression':
        return pyls.pls_regression(X.copy(), Y.copy(),
                                   n_components=w['L'], **kw)
    if w['kind'] == 'meancentered':
        return pyls.meancentered_pls(X, groups=w['groups'],
                                     n_cond=w['n_cond'], mean_centering=0,
                                     n_split=0, rotate=True, ci=95, **kw)
    return pyls.behavioral_pls(X, Y, groups=w['groups'], n_cond=w['n_cond'],
                               n_split=0, test_split=0, rotate=True, ci=95,
                               covariance=False, **kw)


def flatten(res):
    out = {}
    for k in ('x_weights', 'y_weights', 'singvals', 'varexp'):
        v = res.get(k)
        if isinstance(v, np.ndarray):
            out[k] = v
    for k in ('pvals', 'permsamples', 'perm_singval'):
        v = res['permres'].get(k)
        if isinstance(v, np.ndarray):
            out[k] = v
    for k in ('x_weights_normed', 'x_weights_stderr', 'bootsamples',
              'y_loadings_boot', 'y_loadings_ci', 'contrast_boot',
              'contrast_ci'):
        v = res['bootres'].get(k)
        if isinstance(v, np.ndarray):
            out[k] = v
    return out


def main(args):
    from bench import WORKLOADS, make_data
    w = WORKLOADS[args.workload]
    X, Y = make_data(w)
    pyls = load_reference()
    n_proc = max(1, args.n_proc)
    n_each = args.n_each
    calib = None
    if n_each <= 0:
        n0 = max(2 * n_proc, 4)
        t0 = time.perf_counter()
        call(pyls, w, X, Y, n0, n_proc, args.seed)
        dt = time.perf_counter() - t0
        TIMERS.clear()
        n_each = int(min(args.max_each,
                         max(n0, n0 * args.target_s / max(dt, 1e-3))))
        calib = {'n_each': n0, 'seconds': dt}
    for i in range(args.warmup):
        call(pyls, w, X, Y, max(2 * n_proc, n_each // 4), n_proc,
             args.seed + 1000 + i)
    TIMERS.clear()
    step_s, res = [], None
    for i in range(args.steps):
        t0 = time.perf_counter()
        res = call(pyls, w, X, Y, n_each, n_proc, args.seed + i)
        step_s.append(time.perf_counter() - t0)
    total = sum(step_s)
    loops = TIMERS.get('permutation', 0.0) + TIMERS.get('bootstrap', 0.0)
    index = TIMERS.get('gen_permsamp', 0.0) + TIMERS.get('gen_bootsamp', 0.0)
    line = {
        'kind': 'reference', 'workload': args.workload, 'n_each': n_each,
        'n_proc': n_proc, 'blas_threads': args.blas_threads or 'default',
        'steps': args.steps, 'warmup': args.warmup, 'step_s': step_s,
        'resamples_per_s': 2.0 * n_each * args.steps / total,
        # the two resampling loops alone (index generation taken out)
        'resamples_per_s_loops': 2.0 * n_each * args.steps /
        max(loops - index, 1e-9),
        'index_generation_s_per_step': index / max(args.steps, 1),
        'permutation_s_per_step':
        TIMERS.get('permutation', 0.0) / max(args.steps, 1),
        'bootstrap_s_per_step':
        TIMERS.get('bootstrap', 0.0) / max(args.steps, 1),
        'host_cores': len(os.sched_getaffinity(0)), 'calibration': calib,
    }
    if args.save and res is not None:
        np.savez(args.save, **flatten(res))
    print(json.dumps(line))


if __name__ == '__main__':
    main(ARGS)
