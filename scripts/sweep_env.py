"""bench.py device leg under several environment settings (one line each).
Usage: python scripts/sweep_env.py [--workload cfg5] "" "A=1" "A=2 B=3" ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
workload = 'cfg5'
if args and args[0] == '--workload':
    workload, args = args[1], args[2:]
for setting in args or ['']:
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split('=', 1)
        env[k] = v
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--workload', workload,
                        '--steps', '2', '--warmup', '3', '--no-cpu-baseline', '--no-e2e',
                        '--no-parity', '--no-fast-path'], env=env, capture_output=True, text=True)
    line = None
    for l in p.stdout.splitlines():
        if l.startswith('{'):
            line = json.loads(l)
    if line is None:
        print('%-40s FAILED: %s' % (setting, p.stderr[-300:]))
        continue
    ks = ' '.join('%s=%.2f' % (k, v) for k, v in line['kernel_ms_per_step'].items())
    print('%-8s %-40s step=%.2f | %s' % (workload, setting or '(default)', line['ms_per_step'], ks), flush=True)
