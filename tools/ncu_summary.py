"""Summarise ncu outputs into small text files for profiles/.

  python tools/ncu_summary.py launches <launches.csv> > profiles/<name>_launches.txt
  python tools/ncu_summary.py full <prof.ncu-rep>     > profiles/<name>_full.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']


def launches(path):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'), hdr.index('Block Size')
    agg = defaultdict(lambda: [0, 0.0, '', ''])
    for r in rows[1:]:
        a = agg[r[ki]]
        a[0] += 1
        a[1] += float(r[vi].replace(',', ''))
        a[2], a[3] = r[gi], r[bi]
    tot = sum(v[1] for v in agg.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)')
    print('%-72s %5s %12s %7s  %s' % ('kernel', 'n', 'total ms', 'share', 'grid x block (last)'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-72s %5d %12.3f %6.1f%%  %s x %s' % (k[:72], v[0], v[1] / 1e6, 100 * v[1] / tot, v[2], v[3]))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')]
        print('## %s' % name)
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print('%-90s %-14s %s' % (key, units[i], vals[i]))
        print()


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
