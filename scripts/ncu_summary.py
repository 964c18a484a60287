#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: launch list shares and key counters of a full capture.
usage: ncu_summary.py launches <csv> | full <ncu-rep>"""
import collections, csv, subprocess, sys

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split('(')[0].replace('void ', '')
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', ''))
    tot = sum(a[1] for a in agg.values())
    print("%-28s %6s %14s %12s %7s" % ("kernel", "n", "total_ns", "avg_ns", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-28s %6d %14.0f %12.0f %7.3f" % (k, a[0], a[1], a[1] / a[0], a[1] / tot))

KEEP = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct')

def full(path):
    out = subprocess.check_output(['ncu', '-i', path, '--page', 'raw', '--csv'], stderr=subprocess.DEVNULL, text=True)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel =", r[hdr.index('Kernel Name')].split('(')[0])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k); print("  %s = %s %s" % (k, r[i], units[i]))
        print()

if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
