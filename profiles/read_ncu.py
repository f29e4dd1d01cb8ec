"""Print the metrics this repo's roofline discussion uses from an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']
STALL = 'smsp__average_warps_issue_stalled_'


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index('Kernel Name')
    print('kernels:', sorted({r[ki] for r in data}))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f'{k} [{units[i]}]:', ', '.join(r[i] for r in data))
    print('warp stall reasons (warps stalled per issue-active cycle):')
    for i, h in enumerate(hdr):
        if h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            vals = [float(r[i]) for r in data]
            if max(vals) >= 0.1:
                print('  ' + h[len(STALL):-len('_per_issue_active.ratio')] + ':', ', '.join(f'{v:.2f}' for v in vals))


if __name__ == '__main__':
    main(sys.argv[1])
