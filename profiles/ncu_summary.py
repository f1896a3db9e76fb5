#!/usr/bin/env python
"""One-screen summary of every kernel in an ncu report: python profiles/ncu_summary.py X.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum']


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(d['Kernel Name'][:60])
        for w in WANT:
            if w in d:
                print(f'   {w} = {d[w]} {units[hdr.index(w)]}')
        st = [(k, v) for k, v in d.items() if 'issue_stalled' in k and k.endswith('_per_warp_active.pct')]
        for k, v in sorted(st, key=lambda kv: -float(kv[1] or 0))[:6]:
            print(f'      {k.split("issue_stalled_")[1]} {v}')


if __name__ == '__main__':
    main(sys.argv[1])
