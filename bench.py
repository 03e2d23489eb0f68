# -*- coding: utf-8 -*-
"""
Benchmark of the PLS resampling hot path (BASELINE.json: resamples/sec,
permutations + bootstraps combined, on synthetic fp64 X(200x100000) Y(200x10)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload cfg5|cfg2|cfg3|cfg4] [--scaling strong|weak]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (default): BASELINE.json configs[4], the configuration north_star
quotes the metric on -- behavioral_pls X(200x100000) Y(200x10), n_perm=10000,
n_boot=10000.  It fits one GPU; with --gpus N the FIXED 10000 + 10000
resamples are sharded over the N ranks by resample id (`scaling: "strong"`).
cfg2 / cfg3 / cfg4 are the other BASELINE configurations (parity-test cases,
selectable here for profiling; `--scaling weak` runs the workload per GPU).

A "step" is one pass of the hot path over the batch: this rank's block of the
permutations and bootstraps, the closing collectives (all-gather of the
permuted singular values, all-reduce of the bootstrap accumulators, the
series-sharded percentile exchange) and the p-value / bootstrap-ratio /
percentile reductions.  `value` is timed with CUDA events (max over ranks)
with X, Y and the resampling tables resident in HBM; `e2e` times the public
front-end call (pypyls_b200.behavioral_pls) from pinned host arrays to host
results, host<->device copies and on-device index generation included.

CPU side (rank 0, N = 1): the UNMODIFIED reference package (oracle/_ref,
materialised by oracle/build_ref.py; `kind: "reference"`) is driven through
its stock path `pyls.behavioral_pls(..., n_proc=<cores>)` on N_cpu
permutations + N_cpu bootstraps of the same workload, in the three thread
settings SURVEY 8(d) names; its outputs are compared element-wise with the
CUDA path run on the reference's own resampling tables (`parity`), and the
p-values / CIs of the full-n run are compared with the vectorised fast oracle
(oracle/pls_oracle.py::fast_stats, itself held to the reference on the N_cpu
subset).  Falls back to the NumPy oracle port (`kind: "port"`) where
oracle/_ref is absent.

`--impl reference` times that same stock reference path on all host cores.

`roofline` describes the dominant kernel class (the cross-covariance contraction).
By default it runs as int8 digit-plane products on the tcgen05 tensor cores
(csrc/gemm_i8.cu): `achieved` = int8 operations executed (counted by the library,
plsb_gemm_work) / summed CUDA-event time of the class, `peak` = 2 x the measured
bf16 tensor figure of MEASURED_PEAKS.json, plus the FP64-equivalent TFLOP/s against
the cuBLAS DGEMM peak measured in the same run.  `--gemm-backend dmma` runs and
accounts the FP64 DMMA kernel instead (algorithmic flop / DGEMM peak).  For N > 1
the end-to-end call uploads X, Y on rank 0 only and broadcasts them over NCCL
(`input_source="root"`).
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[4]: the configuration the metric is quoted on
    # (north_star: "X(200x100000) Y(200x10) reported at 1/2/4/8 B200")
    'cfg5': dict(kind='behavioral', S=200, B=100000, T=10, groups=[200],
                 n_cond=1, n_perm=10000, n_boot=10000,
                 name='behavioral_pls X(200x100000) Y(200x10) n_perm=10000 '
                      'n_boot=10000'),
    # BASELINE.json configs[1]
    'cfg2': dict(kind='behavioral', S=80, B=10000, T=10, groups=[20, 20],
                 n_cond=2, n_perm=5000, n_boot=5000,
                 name='behavioral_pls X(80x10000) Y(80x10) groups=[20,20] '
                      'n_cond=2 n_perm=5000 n_boot=5000'),
    # BASELINE.json configs[2]; X has 320 rows (the 160 of BASELINE.json do not
    # fit groups=[40]*4 x n_cond=2)
    'cfg3': dict(kind='meancentered', S=320, B=20000, T=1,
                 groups=[40, 40, 40, 40], n_cond=2, n_perm=10000, n_boot=10000,
                 name='meancentered_pls X(320x20000) groups=[40,40,40,40] '
                      'n_cond=2 n_perm=10000 n_boot=10000'),
    # BASELINE.json configs[3]
    'cfg4': dict(kind='regression', S=500, B=5000, T=20, L=10, groups=[500],
                 n_cond=1, n_perm=5000, n_boot=10000,
                 name='pls_regression (SIMPLS) X(500x5000) Y(500x20) '
                      'n_components=10 n_perm=5000 n_boot=10000'),
}
DEFAULT_WORKLOAD = 'cfg5'
METRIC = 'resamples/sec (perm+boot)'


def make_data(w):
    rs = np.random.RandomState(1234)
    X = rs.rand(w['S'], w['B'])
    Y = rs.rand(w['S'], w['T'])
    return X, Y


def host_cores():
    return len(os.sched_getaffinity(0))


# ---------------------------------------------------------------------------
# CPU arm 1: the stock reference package (oracle/_ref) in a subprocess
# ---------------------------------------------------------------------------
def reference_available():
    return os.path.isfile(os.path.join(ROOT, 'oracle', '_ref', 'pyls',
                                       '__init__.py'))


def run_reference(workload, n_each=0, n_proc=1, blas_threads=1, steps=1,
                  warmup=0, target_s=6.0, save=None, timeout=1500):
    """One oracle/ref_runner.py process; returns its JSON line (dict)."""
    cmd = [sys.executable, os.path.join(ROOT, 'oracle', 'ref_runner.py'),
           '--workload', workload, '--n-each', str(n_each), '--n-proc',
           str(n_proc), '--blas-threads', str(blas_threads), '--steps',
           str(steps), '--warmup', str(warmup), '--target-s', str(target_s)]
    if save:
        cmd += ['--save', save]
    env = dict(os.environ)
    for var in ('OPENBLAS_NUM_THREADS', 'OMP_NUM_THREADS', 'MKL_NUM_THREADS'):
        env.pop(var, None)
    p = subprocess.run(cmd, capture_output=True, text=True, env=env,
                       timeout=timeout, cwd=ROOT)
    for line in reversed(p.stdout.splitlines()):
        if line.startswith('{'):
            return json.loads(line)
    raise RuntimeError('oracle/ref_runner.py failed: ' + p.stderr[-800:])


# ---------------------------------------------------------------------------
# CPU arm 2 (second number / fallback): the NumPy oracle port, one process per
# core, BLAS pinned to one thread
# ---------------------------------------------------------------------------
_CPU = {}


def _cpu_init(wname):
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    os.environ['OMP_NUM_THREADS'] = '1'
    try:
        from threadpoolctl import threadpool_limits
        _CPU['limit'] = threadpool_limits(1)
    except Exception:
        pass
    import warnings
    warnings.filterwarnings('ignore')
    from oracle import pls_oracle as po
    w = WORKLOADS[wname]
    X, Y = make_data(w)
    if w['kind'] == 'regression':
        X, Y = X - X.mean(0), Y - Y.mean(0)
        spec = po._Spec('regression', [w['S']], 1, n_components=w['L'])
    elif w['kind'] == 'meancentered':
        spec = po._Spec('meancentered', w['groups'], w['n_cond'])
        Y = spec.dummy
    else:
        spec = po._Spec('behavioral', w['groups'], w['n_cond'])
    U, d, V = po.decompose(spec, X, Y, seed=1234)
    _CPU.update(po=po, X=X, Y=Y, spec=spec, U=U, V=V)


def _cpu_task(args):
    kind, cols = args
    po, c = _CPU['po'], _CPU
    if kind == 'perm':
        po.run_perms(c['spec'], c['X'], c['Y'], cols, c['V'])
    else:
        po.run_boots(c['spec'], c['X'], c['Y'], cols, c['U'])
    return cols.shape[1]


class CpuPort:
    """Times the oracle port's permutation + bootstrap loops on `cores`
    worker processes over a bounded sample of the workload's resamples."""

    def __init__(self, wname, cores=None):
        import multiprocessing as mp
        self.w = WORKLOADS[wname]
        self.cores = cores or host_cores()
        from oracle import pls_oracle as po
        import warnings
        warnings.filterwarnings('ignore')
        n = max(4 * self.cores, 32)
        self.perm = po.gen_permsamp(self.w['groups'], self.w['n_cond'], n,
                                    seed=1)
        self.boot = po.gen_bootsamp(self.w['groups'], self.w['n_cond'], n,
                                    seed=2)
        self.pool = mp.get_context('spawn').Pool(
            self.cores, initializer=_cpu_init, initargs=(wname,))
        # make sure every worker is up and has its data before timing
        self.pool.map(_cpu_task, [('perm', self.perm[:, :1])] * self.cores)

    def step(self, n_each):
        """Runs n_each permutations and n_each bootstraps; returns seconds."""
        per = max(1, n_each // (2 * self.cores))
        tasks = []
        for kind, tab in (('perm', self.perm), ('boot', self.boot)):
            for a in range(0, n_each, per):
                cols = tab[:, [i % tab.shape[1] for i in
                               range(a, min(a + per, n_each))]]
                tasks.append((kind, cols))
        t0 = time.perf_counter()
        done = sum(self.pool.map(_cpu_task, tasks, chunksize=1))
        dt = time.perf_counter() - t0
        assert done == 2 * n_each
        return dt

    def calibrate(self, target_s):
        """Resamples of each kind per step so that a step takes ~target_s."""
        n0 = 2 * self.cores
        dt = self.step(n0)
        rate = 2 * n0 / dt
        return int(max(n0, min(rate * target_s / 2, 20000)))

    def close(self):
        self.pool.close()
        self.pool.join()


def port_rate(wname, target_s):
    arm = CpuPort(wname)
    n_each = arm.calibrate(target_s)
    dt = arm.step(n_each)
    arm.close()
    return {'value': 2 * n_each / dt, 'unit': 'resamples/s',
            'cores': arm.cores, 'kind': 'port',
            'sample': '%d permutations + %d bootstraps of the workload, %d '
                      'worker processes with single-threaded BLAS, NumPy '
                      'restatement oracle/pls_oracle.py (skips check_X_y, '
                      'joblib)' % (n_each, n_each, arm.cores)}


def cpu_baseline_leg(wname, n_cpu, quick=False):
    """Reference CPU run of the same workload in the three thread settings of
    SURVEY 8(d); returns (cpu_baseline dict, path of the saved outputs of the
    n_proc run or None)."""
    cores = host_cores()
    if not reference_available():
        out = port_rate(wname, 10.0)
        out['note'] = 'oracle/_ref absent: NumPy oracle port timed instead'
        return out, None
    save = os.path.join(tempfile.mkdtemp(prefix='plsb_ref_'), 'ref.npz')
    t0 = time.perf_counter()
    par = run_reference(wname, n_each=n_cpu, n_proc=cores, blas_threads=1,
                        save=save)
    settings = {'n_proc=%d, BLAS 1 thread' % cores: par}
    if not quick:
        n_ser = 3 if WORKLOADS[wname]['S'] * WORKLOADS[wname]['B'] > 5e6 else 12
        settings['serial, BLAS 1 thread'] = run_reference(
            wname, n_each=n_ser, n_proc=1, blas_threads=1)
        settings['serial, BLAS default threads'] = run_reference(
            wname, n_each=n_ser, n_proc=1, blas_threads=0)
    brief = {k: {'resamples_per_s': v['resamples_per_s'],
                 'resamples_per_s_loops_only': v['resamples_per_s_loops'],
                 'n_perm': v['n_each'], 'n_boot': v['n_each'],
                 'index_generation_s': v['index_generation_s_per_step'],
                 'seconds': sum(v['step_s'])}
             for k, v in settings.items()}
    sample = ('pyls.behavioral_pls stock path (oracle/_ref, unmodified '
              'reference) on %d permutations + %d bootstraps of the workload, '
              'n_proc=%d joblib workers, BLAS 1 thread (the reference CI\'s '
              'setting); whole call timed (original decomposition and '
              'index generation included)' % (n_cpu, n_cpu, cores))
    out = {'value': par['resamples_per_s'], 'unit': 'resamples/s',
           'cores': cores, 'kind': 'reference', 'sample': sample,
           'n_cpu': n_cpu, 'settings': brief,
           'index_generation_s': par['index_generation_s_per_step'],
           'leg_seconds': None}
    if not quick:
        try:
            out['port'] = port_rate(wname, 4.0)
        except Exception as e:       # the second number is optional
            out['port'] = {'error': str(e)[:200]}
    out['leg_seconds'] = time.perf_counter() - t0
    return out, save


# ---------------------------------------------------------------------------
class ClockSampler:
    """One long-running `nvidia-smi -lms 200` that logs clocks / throttle
    reasons while the timed regions run (B200_PROFILING.md's clocks line)."""

    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.log = tempfile.NamedTemporaryFile('w+', suffix='.csv')
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(device_index),
                 '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                 '-lms', '200'], stdout=self.log, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.log.seek(0)
        out = []
        for line in self.log.read().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) >= 8:
                out.append(f)
        return out


def summarise_clocks(samples):
    if not samples:
        return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
    sm = sorted(float(s[0]) for s in samples)
    reasons = set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
             'sw_power_cap']
    for s in samples:
        for name, v in zip(names, s[4:8]):
            if v.lower().startswith('active'):
                reasons.add(name)
    return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(samples[0][1]),
            'reasons': sorted(reasons), 'samples': len(samples)}


def measure_fp64_peak(torch, device):
    """cuBLAS DGEMM 8192^3, best of 5 (burst) -- MEASURED_PEAKS.json carries no
    FP64 figure, so the roofline denominator is measured in the same run."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = float('inf')
    for _ in range(5):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def algorithmic_work(w, n_perm, n_boot):
    """Algorithmic flop per step and per kernel class."""
    S, B, T = w['S'], w['B'], w['T']
    J = len(w['groups']) * w['n_cond']
    K = J * T
    # 2*S*B*T_eff per resample (T_eff = T behavioural, J mean-centred;
    # SIMPLS: Cov plus t = X0 r, p = X0^T t per component)
    t_eff = {'behavioral': T, 'meancentered': J,
             'regression': T + 2 * w.get('L', 0)}[w['kind']]
    xcov = 2.0 * S * B * t_eff * (n_perm + n_boot)
    return {
        'xcov_gemm': ('tensor', xcov),
        'gram_proj': ('tensor', 2.0 * K * (2 * K) * B * n_boot),
        'accum_u': ('tensor', 2.0 * K * K * B * n_boot),
        'small_decomp': ('tensor', None),
        'build_operands': ('hbm', None),
        'stats': ('hbm', None),
    }


def ncu_traffic(wname, world, cls):
    """DRAM bytes per step of a kernel class from the committed ncu capture of
    this very command (profiles/r2_ncu_traffic.json, written by
    scripts/ncu_traffic.py from the .ncu-rep): dram__bytes_read.sum +
    dram__bytes_write.sum over the class's launches of one step."""
    try:
        tab = json.load(open(os.path.join(ROOT, 'profiles',
                                          'r2_ncu_traffic.json')))
        ent = tab['%s_n%d' % (wname, world)]
        return ent['classes'][cls]['dram_bytes_per_step'], ent['source']
    except Exception:
        return None, None


# ---------------------------------------------------------------------------
def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wname = args.workload
    w = WORKLOADS[wname]
    cores = host_cores()
    budget = 150.0
    target = min(20.0, budget / max(1, args.steps + args.warmup))
    if reference_available():
        r = run_reference(wname, n_each=0, n_proc=cores, blas_threads=1,
                          steps=args.steps, warmup=args.warmup,
                          target_s=target)
        n_each, total = r['n_each'], sum(r['step_s'])
        value = r['resamples_per_s']
        kind = 'reference'
        sample = ('%d permutations + %d bootstraps of the workload per step '
                  'through pyls.%s(..., n_proc=%d) of the unmodified '
                  'reference package (oracle/_ref), BLAS 1 thread per worker; '
                  'whole call timed' % (
                      n_each, n_each,
                      {'behavioral': 'behavioral_pls',
                       'meancentered': 'meancentered_pls',
                       'regression': 'pls_regression'}[w['kind']], cores))
        extra = {'resamples_per_s_loops_only': r['resamples_per_s_loops'],
                 'index_generation_s_per_step':
                 r['index_generation_s_per_step']}
    else:
        arm = CpuPort(wname)
        n_each = arm.calibrate(target_s=target)
        for _ in range(args.warmup):
            arm.step(max(2 * arm.cores, n_each // 4))
        times = [arm.step(n_each) for _ in range(args.steps)]
        arm.close()
        total = sum(times)
        value = 2 * n_each * args.steps / total
        kind = 'port'
        sample = ('%d permutations + %d bootstraps of the workload per step, '
                  'NumPy oracle port, %d worker processes, BLAS 1 thread each '
                  '(oracle/_ref absent)' % (n_each, n_each, arm.cores))
        extra = {}
    line = {
        'impl': 'reference', 'metric': METRIC,
        'value': value, 'unit': 'resamples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic (RandomState(1234).rand)',
        'config': {'workload': w['name'], 'sample': sample},
        'cpu_baseline': dict({'value': value, 'unit': 'resamples/s',
                              'cores': cores, 'kind': kind,
                              'sample': sample}, **extra),
        'e2e': {'value': value, 'unit': 'resamples/s',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# parity
# ---------------------------------------------------------------------------
def _rel(a, b):
    """max |a - b| / max |b| (entries of a distribution can be ~0)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = np.max(np.abs(b)) or 1.0
    return float(np.max(np.abs(a - b)) / scale)


def frontend_call(pyls, w, X, Y, **kw):
    kind = w['kind']
    if kind == 'regression':
        return pyls.pls_regression(X, Y, n_components=w['L'], **kw)
    if kind == 'meancentered':
        return pyls.meancentered_pls(X, groups=w['groups'],
                                     n_cond=w['n_cond'], **kw)
    return pyls.behavioral_pls(X, Y, groups=w['groups'], n_cond=w['n_cond'],
                               **kw)


def parity_vs_reference(pyls, w, X, Y, ref_file, device):
    """Element-wise comparison of the CUDA path with the reference's own
    outputs on the reference's own resampling tables (first N_cpu resamples of
    the workload)."""
    ref = dict(np.load(ref_file))
    n = int(ref['permsamples'].shape[1])
    out = frontend_call(pyls, w, X.copy(), Y.copy(), n_perm=n, n_boot=n,
                        seed=1234, verbose=False, device=device,
                        permsamples=ref['permsamples'],
                        bootsamples=ref['bootsamples'])
    sv_key = 'singvals' if 'singvals' in ref else 'varexp'   # regression: pctvar
    sv_ref = np.asarray(ref[sv_key])
    sv_ref = np.diag(sv_ref) if sv_ref.ndim == 2 else sv_ref
    live = sv_ref > 1e-10 * sv_ref.max()      # numerically null LVs: noise
    boot_key = 'contrast_boot' if w['kind'] == 'meancentered' \
        else 'y_loadings_boot'
    ci_key = 'contrast_ci' if w['kind'] == 'meancentered' else 'y_loadings_ci'
    res = {
        'n_perm': n, 'n_boot': n,
        'singvals_max_rel': _rel(np.asarray(out[sv_key])[live], sv_ref[live]),
        'perm_singval_max_rel': _rel(out.permres.perm_singval[live],
                                     ref['perm_singval'][live]),
        'pvals_max_abs_diff': float(np.max(np.abs(
            out.permres.pvals[live] - ref['pvals'][live]))),
        'distrib_max_rel': _rel(out.bootres[boot_key][:, live],
                                ref[boot_key][:, live]),
        'ci_max_rel': _rel(out.bootres[ci_key][:, live],
                           ref[ci_key][:, live]),
        'ci_max_abs_diff': float(np.max(np.abs(
            out.bootres[ci_key][:, live] - ref[ci_key][:, live]))),
        'bootstrap_ratio_max_rel': _rel(
            out.bootres.x_weights_normed[:, live],
            ref['x_weights_normed'][:, live]),
    }
    res['max_rel'] = max(res['singvals_max_rel'],
                         res['perm_singval_max_rel'], res['distrib_max_rel'],
                         res['ci_max_rel'])
    res['what'] = ('CUDA front-end vs the unmodified reference on the '
                   'reference\'s own %d + %d resampling tables, element-wise; '
                   '*_max_rel = max|a-b| / max|b|' % (n, n))
    return res


def parity_full_n(w, X, Y, out):
    """p-values and CIs of the full-n run against the vectorised fast oracle
    on the SAME (device-generated) tables."""
    from oracle import pls_oracle as po
    if w['kind'] == 'regression':
        return None
    t0 = time.perf_counter()
    if w['kind'] == 'meancentered':
        spec = po._Spec('meancentered', w['groups'], w['n_cond'])
        Yy, boot_key, ci_key = spec.dummy, 'contrast_boot', 'contrast_ci'
    else:
        spec = po._Spec('behavioral', w['groups'], w['n_cond'])
        Yy, boot_key, ci_key = Y, 'y_loadings_boot', 'y_loadings_ci'
    sv = np.asarray(out.singvals)
    sv = np.diag(sv) if sv.ndim == 2 else sv
    live = sv > 1e-10 * sv.max()
    fast = po.fast_stats(spec, X, Yy, out.permres.permsamples,
                         out.bootres.bootsamples, out.x_weights, sv,
                         out.y_weights)
    return {
        'n_perm': int(out.permres.permsamples.shape[1]),
        'n_boot': int(out.bootres.bootsamples.shape[1]),
        'pvals_max_abs_diff': float(np.max(np.abs(
            out.permres.pvals[live] - fast['pvals'][live]))),
        'perm_singval_max_rel': _rel(out.permres.perm_singval[live],
                                     fast['perm_singval'][live]),
        'ci_max_abs_diff': float(np.max(np.abs(
            out.bootres[ci_key][:, live] - fast['distrib_ci'][:, live]))),
        'distrib_max_rel': _rel(out.bootres[boot_key][:, live],
                                fast['distrib'][:, live]),
        'oracle_seconds': time.perf_counter() - t0,
        'what': 'full-n front-end run (device-generated tables) vs '
                'oracle.fast_stats on the same tables',
    }


def rank_invariance(eng, w, out, world, P, R):
    """N > 1: rank 0 recomputes, on its own GPU, the first resamples of EVERY
    rank's block from the tables of the sharded run and compares them with
    what the owning rank produced: the bootstrap distribution must agree bit
    for bit, the permuted singular values to 1e-12 (their row sums of squares
    are reduced over a batch-size dependent number of column splits)."""
    from pypyls_b200.resample import shard_range
    if w['kind'] == 'regression':
        return None
    boot_key = 'contrast_boot' if w['kind'] == 'meancentered' \
        else 'y_loadings_boot'
    n_each, bit_equal, worst = 16, True, 0.0
    for r in range(world):
        fp, cp = shard_range(P, r, world)
        ids = np.arange(fp, fp + min(n_each, cp))
        mine = eng.run_perms(out.permres.permsamples[:, ids],
                             rotate=True).cpu().numpy().T
        theirs = out.permres.perm_singval[:, ids]
        live = theirs.max(axis=1) > 1e-10 * theirs.max()
        worst = max(worst, _rel(mine[live], theirs[live]))
        fb, cb = shard_range(R, r, world)
        ids = np.arange(fb, fb + min(n_each, cb))
        mine = eng.boot_distrib(out.bootres.bootsamples[:, ids])
        mine = mine.cpu().numpy().transpose(1, 2, 0)
        bit_equal = bit_equal and np.array_equal(mine,
                                                 out.bootres[boot_key][..., ids])
    return {'ranks_checked': world, 'resamples_per_rank': n_each,
            'distrib_bit_equal': bool(bit_equal),
            'perm_singval_max_rel': worst,
            'what': 'rank 0 recomputed the first resamples of every rank\'s '
                    'block from the sharded run\'s tables'}


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD,
                    choices=sorted(WORKLOADS))
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'],
                    help='strong: the workload\'s resamples are sharded over '
                         'the ranks; weak: every rank runs the whole workload')
    ap.add_argument('--n-cpu', type=int, default=200,
                    help='permutations and bootstraps of the CPU reference leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--quick-cpu', action='store_true',
                    help='CPU leg: only the n_proc setting')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-fast-path', action='store_true',
                    help='skip the separately reported sample-space permutation leg')
    ap.add_argument('--gemm-backend', default='auto', choices=['auto', 'dmma'],
                    help="cross-covariance contraction: 'auto' = int8 slice GEMM "
                         "on the tcgen05 tensor cores where it applies, 'dmma' = "
                         "FP64 DMMA everywhere")
    ap.add_argument('--gemm-slices', type=int, default=6, choices=[5, 6, 7],
                    help='int8 digit planes per operand of the slice GEMM')
    ap.add_argument('--workspace-gib', type=float, default=None,
                    help='chunk workspace of the engine (default: library default)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == 'reference':
        run_reference_arm(args)
        return

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    wname = args.workload
    w = WORKLOADS[wname]

    # CPU reference leg first (rank 0, N = 1 only), before CUDA is initialised
    cpu_baseline, ref_file = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = args.n_cpu
        cores = host_cores()
        if w['S'] * w['B'] > 5e6 and cores < 16:
            n_cpu = max(32, n_cpu * cores // 16)      # keep the leg bounded
        cpu_baseline, ref_file = cpu_baseline_leg(wname, n_cpu,
                                                  quick=args.quick_cpu)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device'
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    import pypyls_b200 as pyls
    from pypyls_b200 import dist as pdist
    from pypyls_b200.engine import ResamplingEngine
    from pypyls_b200.resample import shard_range

    X, Y = make_data(w)
    Xh = torch.from_numpy(X).pin_memory()
    Yh = torch.from_numpy(Y).pin_memory()
    if args.scaling == 'strong':
        P, R = w['n_perm'], w['n_boot']                 # whole job, fixed
    else:
        P, R = w['n_perm'] * world, w['n_boot'] * world
    first_p, n_perm = shard_range(P, rank, world)
    first_b, n_boot = shard_range(R, rank, world)

    kind = w['kind']
    ws_bytes = None if args.workspace_gib is None else \
        int(args.workspace_gib * (1 << 30))
    if kind == 'regression':
        from pypyls_b200.types.regression import gaussian_tables
        eng = ResamplingEngine('regression', w['S'], w['B'], w['T'], [w['S']],
                               1, device=local_rank, n_components=w['L'],
                               workspace_bytes=ws_bytes,
                               gemm_backend=args.gemm_backend,
                               gemm_slices=args.gemm_slices)
        eng.set_data(Xh - Xh.mean(0, keepdim=True),
                     Yh - Yh.mean(0, keepdim=True))
        rs0 = np.random.RandomState(1234)
        om0 = np.stack([rs0.normal(size=(w['T'], 11)) for _ in range(w['L'])])
        U, d = eng.simpls_decompose(om0 if w['T'] > 11 else None)
        om_p = om_b = None
        if w['T'] > 11:
            om_p = eng.to_device(gaussian_tables(
                range(first_p, first_p + n_perm), w['T']))
            om_b = eng.to_device(gaussian_tables(
                range(first_b, first_b + n_boot), w['T']))
        bs, add_orig = U.contiguous(), True
    else:
        eng = ResamplingEngine(kind, w['S'], w['B'], w['T'], w['groups'],
                               w['n_cond'], device=local_rank,
                               workspace_bytes=ws_bytes,
                               gemm_backend=args.gemm_backend,
                               gemm_slices=args.gemm_slices)
        eng.set_data(Xh, Yh if kind == 'behavioral' else None)
        U, d, V = eng.decompose()
        bs, add_orig = (U * d[None, :]).contiguous(), kind == 'behavioral'

    # on-device index generation, timed on its own
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    idx_p, _ = eng.gen_perm_indices(1234, n_perm, first=first_p)
    idx_b, _ = eng.gen_boot_indices(1234, n_boot, first=first_b)
    torch.cuda.synchronize(device)
    index_gen_s = time.perf_counter() - t0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def finish(d_perm_local, distrib_local, us, uq):
        d_perm = pdist.gather_resamples(d_perm_local, P)
        pdist.reduce_sum(us, uq)
        pv = eng.perm_pvals(d_perm, d)
        bsr, se = eng.boot_ratio(bs, us, uq, R, add_orig)
        lo, hi = pdist.percentile_sharded(eng, distrib_local, R, 2.5, 97.5)
        return pv, bsr, lo, hi

    def step():
        if kind == 'regression':
            d_perm = eng.simpls_run_perms(idx_p, om_p)
            distrib, us, uq, _ = eng.simpls_run_boots(idx_b, om_b)
        else:
            d_perm = eng.run_perms(idx_p, rotate=True)
            distrib, us, uq = eng.run_boots(idx_b)
        return finish(d_perm, distrib, us, uq)

    def step_fast():
        # the same job with the rotated permutations evaluated in sample space
        # (plsb_run_perms_gram): an algorithmic fast path that skips the
        # cross-covariance contraction of the permutations -- reported apart,
        # never divided by the GEMM's algorithmic flop
        d_perm = eng.run_perms_gram(idx_p)
        distrib, us, uq = eng.run_boots(idx_b)
        return finish(d_perm, distrib, us, uq)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        evs = []
        barrier()
        for _ in range(steps):
            flush.fill_(1)                  # evict L2 between timed iterations
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    eng.timing_enable(True)
    eng.timing_read()
    eng.gemm_work(reset=True)
    launches0 = eng.launch_count
    ms = timed(step, args.steps)
    launches = eng.launch_count - launches0
    classes = eng.timing_read()
    i8_macs, dmma_flops = eng.gemm_work(reset=True)
    eng.timing_enable(False)
    value = (P + R) * args.steps / (ms * 1e-3)

    # ---- algorithmic fast path (sample-space permutations), timed the same way --
    fast = None
    if kind != 'regression' and not args.no_fast_path:
        for _ in range(2):
            step_fast()
        fms = timed(step_fast, args.steps)
        fast = {'value': (P + R) * args.steps / (fms * 1e-3),
                'unit': 'resamples/s', 'ms_per_step': fms / args.steps,
                'what': 'same step with the rotated permutations in sample '
                        'space (a^T (X X^T) a through the S x S Gram matrix of '
                        'the data, perm_path="gram"): skips the permutations\' '
                        'cross-covariance contraction, so it is not a GEMM-'
                        'roofline figure'}

    # ---- end to end through the public front-end ---------------------------
    e2e, out = None, None
    if not args.no_e2e:
        def call(seed, Xa=Xh, Ya=Yh):
            kw = dict(n_perm=P, n_boot=R, seed=seed, verbose=False,
                      device=local_rank, workspace_bytes=ws_bytes,
                      gather_results='root', gemm_backend=args.gemm_backend,
                      gemm_slices=args.gemm_slices)
            if world > 1:
                # rank 0 uploads X, Y from its pinned host memory, the other ranks receive
                # their replicas over NCCL (N simultaneous 160 MB uploads compete for the
                # host's memory bandwidth: 14 ms of fixed cost per call at N = 8)
                kw['input_source'] = 'root'
            return frontend_call(pyls, w, Xa, Ya, **kw)
        # warm-up holds on to the previous result like the timed loop does, so that
        # the pinned host blocks of two live results exist before timing starts
        for i in range(3):
            out = call(100 + i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            tc = time.perf_counter()
            out = call(i)
            if os.environ.get('PLSB_BENCH_DEBUG'):
                print('rank %d e2e call %d: %.1f ms' % (
                    rank, i, 1e3 * (time.perf_counter() - tc)), file=sys.stderr)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        d2h = 0
        for res in (out, out.permres, out.bootres):
            for k, v in res.items():
                if isinstance(v, np.ndarray) and k not in ('X', 'Y'):
                    d2h += v.nbytes // (2 if v.dtype == np.int64 else 1)
        e2e = {'value': (P + R) * args.steps / dt, 'unit': 'resamples/s',
               'h2d_bytes_per_step': int(X.nbytes + Y.nbytes),
               'd2h_bytes_per_step': int(d2h),
               'ms_per_step': 1e3 * dt / args.steps,
               'inputs': ('pinned host tensors of rank 0, uploaded once and '
                          'broadcast to the other ranks over NCCL '
                          '(input_source="root")' if world > 1 else
                          'pinned host tensors') +
                         '; results on the host of rank 0 (gather_results="root")'}
        if world == 1:
            # what a user of the reference passes: pageable NumPy arrays
            call(50, X, Y)
            t0 = time.perf_counter()
            for i in range(2):
                call(60 + i, X, Y)
            torch.cuda.synchronize(device)
            e2e['pageable_numpy_ms_per_step'] = \
                1e3 * (time.perf_counter() - t0) / 2
    samples = sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- parity (rank 0) -----------------------------------------------------
    parity = None
    if not args.no_parity:
        parity = {}
        try:
            if ref_file is not None:
                parity['vs_reference'] = parity_vs_reference(
                    pyls, w, X, Y, ref_file, local_rank)
            if out is not None and P + R <= 40000:
                parity['full_n_vs_fast_oracle'] = parity_full_n(w, X, Y, out)
            if out is not None and world > 1:
                parity['rank_invariance'] = rank_invariance(eng, w, out, world,
                                                            P, R)
            rels = [v['max_rel'] for k, v in parity.items()
                    if v and 'max_rel' in v]
            parity['max_rel'] = max(rels) if rels else None
        except Exception as e:          # a parity failure must not hide the timing
            parity['error'] = '%s: %s' % (type(e).__name__, str(e)[:300])

    # ---- roofline of the dominant kernel class -------------------------------
    work = algorithmic_work(w, n_perm, n_boot)
    top = max(classes, key=lambda k: classes[k][0])
    top_ms, top_n = classes[top]
    bound, flops = work.get(top, ('hbm', None))
    peak_tf = measure_fp64_peak(torch, device)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    roofline = {'kernel': top, 'bound': bound, 'achieved': None,
                'peak': None, 'unit': None, 'frac': None, 'traffic': None,
                'kernel_ms_per_step': top_ms / args.steps,
                'launches_per_step': top_n / args.steps}
    if args.workspace_gib is None and args.scaling == 'strong':
        traffic, source = ncu_traffic(wname, world, top)
        if traffic is not None:
            roofline['traffic'] = traffic
            roofline['traffic_unit'] = 'bytes per step, all launches of the ' \
                'class (dram__bytes_read.sum + dram__bytes_write.sum)'
            roofline['traffic_source'] = source
    roofline['algorithmic_flop_per_step'] = flops
    # flop the launches really execute: rotated permutations contract L rows per
    # resample (|R^T v_j| for every original y-weight) instead of T per cell;
    # bootstraps the block-diagonal operand (T rows per cell over the cell's rows)
    if top == 'xcov_gemm' and kind == 'behavioral':
        J = len(w['groups']) * w['n_cond']
        ng = w['S'] / J
        executed = 2.0 * w['S'] * w['B'] * (J * w['T']) * n_perm + \
            2.0 * ng * w['B'] * (J * w['T']) * n_boot
        roofline['executed_flop_per_step'] = executed
    if bound == 'tensor' and flops and top == 'xcov_gemm' and i8_macs > 0:
        # The class ran (mostly) as int8 digit-plane products on the tcgen05
        # tensor cores: its ceiling is the int8 tensor rate.  achieved = int8
        # operations the launches executed (padded tiles x plane products, from the
        # library's own counter) / summed launch time; the FP64 flop the DMMA
        # kernel still executed in the class (squared-operand statistics of grouped
        # layouts, ...) are not counted, so the fraction is a lower bound.
        ach = 2.0 * i8_macs / (top_ms * 1e-3) / 1e12
        if peaks.get('bf16_tflops'):
            peak, src = 2.0 * peaks['bf16_tflops'], \
                '2 x MEASURED_PEAKS.json bf16_tflops (dense int8 issues at twice ' \
                'the bf16 rate on sm_100a; scripts/probes/tc_rate_probe.cu ' \
                'measured 4.2-4.4 POP/s of back-to-back int8 MMAs on this pool)'
        else:
            peak, src = 4500.0, 'fallback: nominal dense int8 4.5 POP/s'
        fp64_eq = flops * args.steps / (top_ms * 1e-3) / 1e12
        roofline.update(achieved=ach, peak=peak, unit='TFLOP/s',
                        unit_detail='int8 tensor operations (TOP/s): achieved and '
                                    'peak count int8 multiply-adds x 2',
                        frac=ach / peak, peak_source=src,
                        gemm_backend='int8 slice GEMM, %d digit planes (%d plane '
                                     'products per FP64 product)' % (
                                         args.gemm_slices, args.gemm_slices *
                                         (args.gemm_slices + 1) // 2),
                        int8_ops_per_step=2.0 * i8_macs / args.steps,
                        dmma_flop_per_step=dmma_flops / args.steps,
                        fp64_equivalent_tflops=fp64_eq,
                        fp64_dgemm_peak_tflops=peak_tf,
                        fp64_equivalent_over_dgemm_peak=fp64_eq / peak_tf)
    elif bound == 'tensor' and flops:
        ach = flops * args.steps / (top_ms * 1e-3) / 1e12
        if roofline.get('executed_flop_per_step'):
            roofline['frac_executed'] = roofline['executed_flop_per_step'] * \
                args.steps / (top_ms * 1e-3) / 1e12 / peak_tf
        roofline.update(achieved=ach, peak=peak_tf, unit='TFLOP/s',
                        frac=ach / peak_tf,
                        peak_source='cuBLAS DGEMM 8192^3 measured in this run '
                                    '(FP64; MEASURED_PEAKS.json has no FP64 '
                                    'figure)')
    else:
        roofline.update(peak=peaks.get('hbm_gbs', 6650.0), unit='GB/s',
                        peak_source='MEASURED_PEAKS.json hbm_gbs'
                        if peaks else 'fallback 6.65 TB/s')
    per_gpu = '' if args.scaling == 'strong' else ' per GPU'
    line = {
        'metric': METRIC, 'value': value,
        'unit': 'resamples/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': warm, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': args.scaling,
        'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic (RandomState(1234).rand)',
        'config': {'workload': w['name'] + per_gpu,
                   'parallelism': 'resample ids sharded over %d GPU(s) '
                                  '(%d permutations + %d bootstraps on this '
                                  'rank); all-gather + all-reduce + series '
                                  'exchange at the end' % (world, n_perm,
                                                           n_boot),
                   'l2': 'flushed between timed steps (256 MiB fill); every '
                         'step also streams a multi-GB cross-covariance '
                         'workspace'},
        'e2e': e2e, 'gpu_launches': int(launches),
        'clocks': summarise_clocks(samples), 'roofline': roofline,
        'cpu_baseline': cpu_baseline, 'parity': parity,
        'index_generation': {
            'device_s': index_gen_s,
            'device_tables': '%d permutation + %d bootstrap columns of this '
                             'rank (columns [0, first + count) are generated '
                             'and de-duplicated)' % (n_perm, n_boot),
            'reference_s': None if not cpu_baseline else
            cpu_baseline.get('index_generation_s'),
            'reference_tables': None if not cpu_baseline else
            '%d + %d columns (gen_permsamp / gen_bootsamp are O(n^2 S))' % (
                cpu_baseline.get('n_cpu', 0), cpu_baseline.get('n_cpu', 0))},
        'kernel_ms_per_step': {k: v[0] / args.steps
                               for k, v in classes.items() if v[1]},
        'fp64_dgemm_tflops_measured': peak_tf,
        'algorithmic_fast_path': fast,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
