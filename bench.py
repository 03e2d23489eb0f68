# -*- coding: utf-8 -*-
"""
Benchmark of the PLS resampling hot path (BASELINE.json: resamples/sec,
permutations + bootstraps combined).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: n_perm permutations and
n_boot bootstraps of the workload (per GPU; weak scaling), including the
p-value / bootstrap-ratio / percentile reductions and, for N > 1, the closing
all-gather + all-reduce.  `value` is timed with CUDA events with X, Y and the
resampling tables resident in HBM; `e2e` times the public front-end call
(pypyls_b200.behavioral_pls) from pinned host arrays to host results, host<->
device copies and on-device index generation included.

`--impl reference` times the CPU implementation of the same path (the NumPy
oracle port of the reference, oracle/pls_oracle.py -- the reference itself is
a Python package that is not present on the GPU box) on all host cores, one
process per core with single-threaded BLAS, which is the reference's own
parallel mode (pyls/utils.py:252-279, .travis.yml:22-23).
"""

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
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
    # BASELINE.json configs[4] (per GPU share is set by --gpus in a real run)
    'cfg5': dict(kind='behavioral', S=200, B=100000, T=10, groups=[200],
                 n_cond=1, n_perm=10000, n_boot=10000,
                 name='behavioral_pls X(200x100000) Y(200x10) n_perm=10000 '
                      'n_boot=10000'),
}


def make_data(w):
    rs = np.random.RandomState(1234)
    X = rs.rand(w['S'], w['B'])
    Y = rs.rand(w['S'], w['T'])
    return X, Y


# ---------------------------------------------------------------------------
# CPU arm (oracle port): one process per core, BLAS pinned to one thread
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


class CpuArm:
    """Times the oracle's permutation + bootstrap loops on `cores` worker
    processes over a bounded sample of the workload's resamples."""

    def __init__(self, wname, cores=None):
        import multiprocessing as mp
        self.w = WORKLOADS[wname]
        self.cores = cores or len(os.sched_getaffinity(0))
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
        n = int(max(n0, min(rate * target_s / 2, 20000)))
        return n

    def close(self):
        self.pool.close()
        self.pool.join()


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
        import tempfile
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
    """Algorithmic flop / bytes per step and per kernel class."""
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


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wname = args.workload
    w = WORKLOADS[wname]
    arm = CpuArm(wname)
    n_each = arm.calibrate(target_s=min(20.0, 150.0 / max(1, args.steps +
                                                          args.warmup)))
    for _ in range(args.warmup):
        arm.step(max(2 * arm.cores, n_each // 4))
    times = [arm.step(n_each) for _ in range(args.steps)]
    arm.close()
    total = sum(times)
    value = 2 * n_each * args.steps / total
    sample = ('%d permutations + %d bootstraps of the workload per step '
              '(first columns of seeded tables), %d worker processes, BLAS 1 '
              'thread each' % (n_each, n_each, arm.cores))
    line = {
        'impl': 'reference', 'metric': 'resamples/sec (perm+boot)',
        'value': value, 'unit': 'resamples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic (RandomState(1234).rand)',
        'config': {'workload': w['name'], 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'resamples/s',
                         'cores': arm.cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'resamples/s',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-fast-path', action='store_true',
                    help='skip the separately reported sample-space permutation leg')
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
    w = WORKLOADS[args.workload]

    # CPU baseline first (rank 0, N = 1 only), before CUDA is initialised
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(args.workload)
        n_each = arm.calibrate(target_s=12.0)
        dt = arm.step(n_each)
        arm.close()
        cpu_baseline = {
            'value': 2 * n_each / dt, 'unit': 'resamples/s',
            'cores': arm.cores, 'kind': 'port',
            'sample': '%d permutations + %d bootstraps of the workload, %d '
                      'worker processes with single-threaded BLAS (the '
                      'reference\'s n_proc mode)' % (n_each, n_each,
                                                     arm.cores)}

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

    X, Y = make_data(w)
    Xh = torch.from_numpy(X).pin_memory()
    Yh = torch.from_numpy(Y).pin_memory()
    n_perm, n_boot = w['n_perm'], w['n_boot']          # per GPU (weak scaling)
    P, R = n_perm * world, n_boot * world               # whole job

    kind = w['kind']
    ws_bytes = None if args.workspace_gib is None else \
        int(args.workspace_gib * (1 << 30))
    if kind == 'regression':
        from pypyls_b200.types.regression import gaussian_tables
        eng = ResamplingEngine('regression', w['S'], w['B'], w['T'], [w['S']],
                               1, device=local_rank, n_components=w['L'],
                               workspace_bytes=ws_bytes)
        eng.set_data(Xh - Xh.mean(0, keepdim=True),
                     Yh - Yh.mean(0, keepdim=True))
        rs0 = np.random.RandomState(1234)
        om0 = np.stack([rs0.normal(size=(w['T'], 11)) for _ in range(w['L'])])
        U, d = eng.simpls_decompose(om0 if w['T'] > 11 else None)
        om_p = om_b = None
        if w['T'] > 11:
            om = eng.to_device(gaussian_tables(
                range(max(n_perm, n_boot) * world), w['T']))
            om_p = om[rank * n_perm:(rank + 1) * n_perm]
            om_b = om[rank * n_boot:(rank + 1) * n_boot]
        bs, add_orig = U.contiguous(), True
    else:
        eng = ResamplingEngine(kind, w['S'], w['B'], w['T'], w['groups'],
                               w['n_cond'], device=local_rank,
                               workspace_bytes=ws_bytes)
        eng.set_data(Xh, Yh if kind == 'behavioral' else None)
        U, d, V = eng.decompose()
        bs, add_orig = (U * d[None, :]).contiguous(), kind == 'behavioral'
    idx_p, _ = eng.gen_perm_indices(1234, n_perm, first=rank * n_perm)
    idx_b, _ = eng.gen_boot_indices(1234, n_boot, first=rank * n_boot)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def step():
        if kind == 'regression':
            d_perm = eng.simpls_run_perms(idx_p, om_p)
            distrib, us, uq, _ = eng.simpls_run_boots(idx_b, om_b)
        else:
            d_perm = eng.run_perms(idx_p, rotate=True)
            distrib, us, uq = eng.run_boots(idx_b)
        if world > 1:
            d_perm = pdist.gather_resamples(d_perm, P)
            distrib = pdist.gather_resamples(distrib, R)
            pdist.reduce_sum(us, uq)
        pv = eng.perm_pvals(d_perm, d)
        bsr, se = eng.boot_ratio(bs, us, uq, R, add_orig)
        lo, hi = eng.percentile(distrib, 2.5, 97.5)
        return pv, bsr, lo, hi

    def step_fast():
        # the same job with the rotated permutations evaluated in sample space
        # (plsb_run_perms_gram): an algorithmic fast path that skips the
        # cross-covariance contraction of the permutations -- reported apart,
        # never divided by the GEMM's algorithmic flop
        d_perm = eng.run_perms_gram(idx_p)
        distrib, us, uq = eng.run_boots(idx_b)
        if world > 1:
            d_perm = pdist.gather_resamples(d_perm, P)
            distrib = pdist.gather_resamples(distrib, R)
            pdist.reduce_sum(us, uq)
        pv = eng.perm_pvals(d_perm, d)
        bsr, se = eng.boot_ratio(bs, us, uq, R, add_orig)
        lo, hi = eng.percentile(distrib, 2.5, 97.5)
        return pv, bsr, lo, hi

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    eng.timing_enable(True)
    eng.timing_read()
    launches0 = eng.launch_count
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)                      # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = eng.launch_count - launches0
    classes = eng.timing_read()
    eng.timing_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = (P + R) * args.steps / (ms * 1e-3)

    # ---- algorithmic fast path (sample-space permutations), timed the same way --
    fast = None
    if kind != 'regression' and not args.no_fast_path:
        for _ in range(2):
            step_fast()
        barrier()
        fevs = []
        for _ in range(args.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            step_fast()
            e1.record()
            fevs.append((e0, e1))
        barrier()
        fms = sum(a.elapsed_time(b) for a, b in fevs)
        ft = torch.tensor([fms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
        fms = float(ft.item())
        fast = {'value': (P + R) * args.steps / (fms * 1e-3),
                'unit': 'resamples/s', 'ms_per_step': fms / args.steps,
                'what': 'same step with the rotated permutations in sample '
                        'space (a^T (X X^T) a through the S x S Gram matrix of '
                        'the data, perm_path="gram"): skips the permutations\' '
                        'cross-covariance contraction, so it is not a GEMM-'
                        'roofline figure'}

    # ---- end to end through the public front-end ---------------------------
    e2e = None
    if not args.no_e2e:
        def call(seed):
            kw = dict(n_perm=P, n_boot=R, seed=seed, verbose=False,
                      device=local_rank, workspace_bytes=ws_bytes)
            if kind == 'regression':
                return pyls.pls_regression(Xh, Yh, n_components=w['L'], **kw)
            if kind == 'meancentered':
                return pyls.meancentered_pls(Xh, groups=w['groups'],
                                             n_cond=w['n_cond'], **kw)
            return pyls.behavioral_pls(Xh, Yh, groups=w['groups'],
                                       n_cond=w['n_cond'], **kw)
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
               'ms_per_step': 1e3 * dt / args.steps}
    samples = sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

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
    # DRAM traffic of the class per step from the committed `ncu --set full`
    # capture (dram__bytes_read.sum + dram__bytes_write.sum over its launches,
    # profiles/r1_ncu_full_v11_summary.csv: permutation launch 0.135 + 0.006 GB,
    # bootstrap launch 1.763 + 16.130 GB -- the stored cross-covariances are 16.2 GB
    # of it); only known for the configuration that was captured
    if (args.workload, world, top) == ('cfg2', 1, 'xcov_gemm') and \
            args.workspace_gib is None:
        roofline['traffic'] = 18.03e9
        roofline['traffic_unit'] = 'bytes per step, all launches of the class'
    roofline['algorithmic_flop_per_step'] = flops
    # flop the launches really execute: rotated permutations contract L rows per
    # resample (|R^T v_j| for every original y-weight) instead of T per cell
    if top == 'xcov_gemm' and kind == 'behavioral':
        J = len(w['groups']) * w['n_cond']
        ng = w['S'] / J
        # bootstraps: the block-diagonal operand (T rows per cell over the cell's
        # ng rows); cells too tall for colstats_kernel's shared-memory tile also
        # run the two count-operand GEMMs of the column statistics
        extra = 0 if ng * 128 * 8 <= 100 * 1024 else 2
        executed = 2.0 * w['S'] * w['B'] * (J * w['T']) * n_perm + \
            2.0 * ng * w['B'] * (J * (w['T'] + extra)) * n_boot
        roofline['executed_flop_per_step'] = executed
    if bound == 'tensor' and flops:
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
    line = {
        'metric': 'resamples/sec (perm+boot)', 'value': value,
        'unit': 'resamples/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic (RandomState(1234).rand)',
        'config': {'workload': w['name'] + ' per GPU, 1xB200 each',
                   'parallelism': 'resamples sharded over %d GPU(s); '
                                  'all-gather + all-reduce at the end' % world,
                   'l2': 'flushed between timed steps (256 MiB fill); every '
                         'step also streams a multi-GB cross-covariance '
                         'workspace'},
        'e2e': e2e, 'gpu_launches': int(launches),
        'clocks': summarise_clocks(samples), 'roofline': roofline,
        'cpu_baseline': cpu_baseline,
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
