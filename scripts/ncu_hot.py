"""Top stall locations of one kernel launch of an ncu report (needs --import-source on).
Usage: python scripts/ncu_hot.py report.ncu-rep launch_index [n_top]"""
import csv, io, subprocess, sys

def main(rep, idx, ntop=40):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(idx),
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    print(rows[0][1][:120])
    hdr, data = rows[1], [r for r in rows[2:] if len(r) > 3]
    si = hdr.index('Warp Stall Sampling (All Samples)')
    def val(r):
        try:
            return int(r[si])
        except ValueError:
            return 0
    tot = sum(val(r) for r in data)
    print('total samples', tot, 'instructions', len(data))
    top = sorted(range(len(data)), key=lambda i: -val(data[i]))[:ntop]
    for i in sorted(top):
        print('%5d %-100s %7d %5.1f%%' % (i, data[i][1].strip()[:100], val(data[i]),
                                          100.0 * val(data[i]) / max(tot, 1)))

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 40)
