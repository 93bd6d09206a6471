"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / bench.py quote.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.avg',
        'smsp__inst_executed_pipe_lsu.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__f_wavefronts.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index('Kernel Name')
    for r in rows[2:]:
        print('=' * 100)
        print(r[name_col][:120])
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"  {h:75s} {units[i]:16s} {r[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
                try:
                    stalls.append((float(r[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        print('  stalls (warps per issue):', ', '.join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:7]))


if __name__ == '__main__':
    main(sys.argv[1])
