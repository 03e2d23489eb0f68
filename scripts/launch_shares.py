"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of one
bench step into per-kernel launch counts, total time and share.

Usage: python scripts/launch_shares.py gpurun_out/launches.csv [first_kernel_substring] > out.md

Only the launches of the LAST timed step are counted: everything from the last
`FillFunctor` launch over the 256 MiB L2-flush buffer (bench.py) onwards, up to the
cuBLAS DGEMM that measures the FP64 peak afterwards."""
import csv
import sys


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        name = r['Kernel Name'].split('(')[0].replace('plsb::<unnamed>::', '')
        name = name.replace('void ', '')
        rows.append((name, float(r['Metric Value']) * 1e-6, r['Grid Size']))
    # last L2 flush = start of the timed step
    start = max((i for i, r in enumerate(rows) if 'FillFunctor' in r[0] and
                 r[2].strip('()').split(',')[0].strip() not in ('1',) and r[1] > 0.02),
                default=0)
    end = len(rows)
    for i in range(start, len(rows)):
        if 'dgemm' in rows[i][0].lower() or 'cutlass' in rows[i][0].lower() or \
                'distribution' in rows[i][0].lower():
            end = i
            break
    agg, order = {}, []
    for name, ms, _ in rows[start + 1:end]:
        if name not in agg:
            agg[name] = [0, 0.0]
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += ms
    total = sum(v[1] for v in agg.values())
    print('| kernel | launches | ms | share |')
    print('|---|---|---|---|')
    for name in order:
        n, ms = agg[name]
        print('| %s | %d | %.3f | %.1f%% |' % (name[:70], n, ms, 100 * ms / total))
    print('| total | | %.3f | |' % total)


if __name__ == '__main__':
    main(sys.argv[1])
