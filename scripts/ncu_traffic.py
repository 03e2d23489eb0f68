"""DRAM traffic and time per kernel class of ONE bench step from an ncu launch list made with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file launches.csv python bench.py --steps 1 ...
Writes / updates profiles/r2_ncu_traffic.json (read by bench.py for `roofline.traffic`) and
prints a markdown table of launch shares.
Usage: python scripts/ncu_traffic.py launches.csv <key, e.g. cfg5_n1> <source note>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = [('xcov_gemm', 'xcov_gemm'), ('gram_proj', 'gram_proj'), ('accum_u', 'accum_u'),
           ('reduce_partials', 'accum_u'), ('build_', 'build_operands'),
           ('ns_rotation', 'small_decomp'), ('eigen_kernel', 'small_decomp'),
           ('rotation_kernel', 'small_decomp'), ('simpls_kernel', 'small_decomp'),
           ('colscale', 'stats'), ('colstats', 'stats'), ('percentile', 'stats'),
           ('pvals', 'stats'), ('boot_ratio', 'stats'), ('finish_rowsq', 'stats'),
           ('transpose', 'prep')]


def classify(name):
    for key, cls in CLASSES:
        if key in name:
            return cls
    return 'other'


def main(path, key, source):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    launches = {}
    order = []
    for r in csv.DictReader(lines):
        i = int(r['ID'])
        if i not in launches:
            name = r['Kernel Name'].split('(')[0].replace('plsb::<unnamed>::', '')
            launches[i] = {'name': name.replace('void ', ''), 'grid': r['Grid Size']}
            order.append(i)
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        m = r['Metric Name']
        if m.startswith('gpu__time_duration'):
            v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
            launches[i]['ms'] = v
        else:
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
            launches[i]['bytes'] = launches[i].get('bytes', 0.0) + v
    rows = [launches[i] for i in order]
    # the timed step = everything after the last 256 MiB L2-flush fill, up to the DGEMM probe
    start = max((i for i, r in enumerate(rows) if 'FillFunctor' in r['name'] and
                 r.get('ms', 0) > 0.02), default=-1)
    end = len(rows)
    for i in range(start + 1, len(rows)):
        n = rows[i]['name'].lower()
        if 'dgemm' in n or 'cutlass' in n or 'distribution' in n or 'gemm_kernel' in n and 'xcov' not in n:
            end = i
            break
    step = rows[start + 1:end]
    agg, per_kernel = {}, {}
    for r in step:
        c = classify(r['name'])
        a = agg.setdefault(c, {'ms': 0.0, 'dram_bytes_per_step': 0.0, 'launches': 0})
        a['ms'] += r.get('ms', 0.0)
        a['dram_bytes_per_step'] += r.get('bytes', 0.0)
        a['launches'] += 1
        k = per_kernel.setdefault(r['name'][:64], [0, 0.0, 0.0])
        k[0] += 1
        k[1] += r.get('ms', 0.0)
        k[2] += r.get('bytes', 0.0)
    total = sum(a['ms'] for a in agg.values())
    print('| kernel | launches | ms | share | DRAM GB |')
    print('|---|---|---|---|---|')
    for name, (n, ms, b) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
        print('| %s | %d | %.3f | %.1f%% | %.2f |' % (name, n, ms, 100 * ms / total, b / 1e9))
    print('| total | | %.3f | | %.2f |' % (total, sum(a['dram_bytes_per_step'] for a in agg.values()) / 1e9))
    out = os.path.join(ROOT, 'profiles', 'r2_ncu_traffic.json')
    tab = json.load(open(out)) if os.path.exists(out) else {}
    tab[key] = {'source': source, 'step_ms_under_ncu': total, 'classes': agg}
    json.dump(tab, open(out, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
