#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, without a GPU) into the handful of metrics the roofline needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [-o profiles/name.md]
"""
import csv
import io
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
    'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct',
]


def main():
    path = sys.argv[1]
    out = sys.argv[sys.argv.index('-o') + 1] if '-o' in sys.argv else None
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    lines = ['# ncu summary of `{}`'.format(path), '']
    name_col = header.index('Kernel Name')
    for k, row in enumerate(data):
        lines.append('## launch {}: `{}` grid {} block {}'.format(
            row[0], row[name_col], row[header.index('Grid Size')], row[header.index('Block Size')]))
        lines.append('')
        lines.append('| metric | value | unit |')
        lines.append('|---|---|---|')
        for want in WANT:
            for i, h in enumerate(header):
                if h == want or h.endswith('.' + want):
                    lines.append('| {} | {} | {} |'.format(want, row[i], units[i]))
                    break
        lines.append('')
    text = '\n'.join(lines)
    if out:
        open(out, 'w').write(text + '\n')
    print(text)


if __name__ == '__main__':
    main()
