"""Warp-stall samples per CUDA source line of one kernel of an ncu report
(needs -lineinfo and --import-source on).
Usage: python scripts/ncu_lines.py report.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main(rep, ntop=40):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source',
                          'sass,cuda'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if r and r[0] == 'Line No')
    si = hdr.index('Warp Stall Sampling (All Samples)')
    agg, src, fname = defaultdict(int), {}, None
    for r in rows:
        if r and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        if r and r[0] == 'Function Name':
            print(r[1][:120])
        if len(r) <= si or not r[0].isdigit():
            continue
        try:
            n = int(r[si] or 0)
        except ValueError:
            continue
        agg[(fname, int(r[0]))] += n
        src[(fname, int(r[0]))] = r[1].strip()
    tot = sum(agg.values())
    print('total samples', tot)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:ntop]:
        print('%-14s %5d %6d %5.1f%%  %s' % (k[0], k[1], v, 100.0 * v / max(tot, 1), src[k][:90]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
