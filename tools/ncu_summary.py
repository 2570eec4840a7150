#!/usr/bin/env python
"""Per-kernel summary of an ncu report: ncu_summary.py REPORT"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d['Kernel Name'][:70])
    for k in want:
        if k in d:
            print('   %-70s %s %s' % (k, d[k], units[hdr.index(k)]))
    st = []
    for k in hdr:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v > 0.25:
                st.append((v, k.split('issue_stalled_')[1].split('_per')[0]))
    print('   stalls/issue:', ', '.join('%s %.2f' % (n, v) for v, n in sorted(st, reverse=True)))
