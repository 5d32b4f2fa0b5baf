#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into a small CSV for profiles/: one row per profiled launch with the metrics the
roofline discussion in DESIGN.md uses.  usage: python tools/ncu_summary.py gpurun_out/full_X.ncu-rep profiles/X_ncu_full.csv"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['%s [%s]' % (n, units[i]) if units[i] else n for n, i in idx])
        for r in rows[2:]:
            w.writerow([r[i].split('(')[0] if n == 'Kernel Name' else r[i] for n, i in idx])
    print('wrote', out, len(rows) - 2, 'launches')


if __name__ == '__main__':
    main()
