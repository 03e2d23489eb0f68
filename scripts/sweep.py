"""Runs bench.py (device-resident leg only) under several environment settings and
prints one compact line of per-kernel-class milliseconds for each.

Usage: python scripts/sweep.py [--workload cfg2] [--steps 2] "" "PLSB_X=1" "PLSB_X=2 PLSB_Y=3" ...
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    args = sys.argv[1:]
    workload, steps = 'cfg2', '2'
    while args and args[0].startswith('--'):
        if args[0] == '--workload':
            workload = args[1]
        elif args[0] == '--steps':
            steps = args[1]
        args = args[2:]
    for setting in args or ['']:
        env = dict(os.environ)
        for kv in setting.split():
            k, v = kv.split('=', 1)
            env[k] = v
        p = subprocess.run(
            [sys.executable, os.path.join(ROOT, 'bench.py'), '--workload', workload,
             '--steps', steps, '--warmup', '3', '--no-cpu-baseline', '--no-e2e'],
            env=env, capture_output=True, text=True)
        line = None
        for l in p.stdout.splitlines():
            if l.startswith('{'):
                line = json.loads(l)
        if line is None:
            print('%-40s FAILED: %s' % (setting, p.stderr[-400:]))
            continue
        ks = ' '.join('%s=%.2f' % (k, v) for k, v in line['kernel_ms_per_step'].items())
        print('%-8s %-40s step=%.2fms value=%.0f | %s' % (
            workload, setting or '(default)', line['ms_per_step'], line['value'], ks), flush=True)


if __name__ == '__main__':
    main()
