"""Warp-stall sample breakdown + a few throughput metrics of one launch of an ncu report.
Usage: python scripts/ncu_stalls.py report.ncu-rep launch_index"""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__cycles_active.avg']

def main(rep, idx):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--launch-skip', str(idx),
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    print(d.get('Kernel Name', '')[:100], d.get('Grid Size'), d.get('Block Size'))
    for k in KEYS:
        for h in hdr:
            if h.startswith(k):
                print('  %-95s %s %s' % (h, d[h], units[hdr.index(h)]))
    st = {h[len('smsp__pcsamp_warps_issue_stalled_'):]: float(d[h]) for h in hdr
          if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')}
    tot = sum(st.values())
    for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
        if v:
            print('  stall %-28s %9d %5.1f%%' % (k, v, 100 * v / tot))

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]))
