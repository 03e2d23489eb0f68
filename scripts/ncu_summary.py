"""Condenses an ncu report (.ncu-rep, `--set full`) into a small CSV of the metrics the
roofline discussion uses.  Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep out.csv"""
import csv, subprocess, sys, io

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'dmma_pipe_pct_active'),
    ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_pct_elapsed'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64_pipe_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__occupancy_limit_shared_mem', 'occ_limit_smem_blocks'),
    ('launch__occupancy_limit_registers', 'occ_limit_reg_blocks'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts'),
    ('smsp__inst_executed.sum', 'warp_instructions'),
]

def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'grid', 'block'] + ['%s [%s]' % (n, units[hdr.index(m)]) for m, n in METRICS if m in hdr])
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')].split('(')[0].replace('plsb::<unnamed>::', '')
            w.writerow([name, r[hdr.index('Grid Size')], r[hdr.index('Block Size')]] +
                       [r[hdr.index(m)] for m, n in METRICS if m in hdr])

if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
