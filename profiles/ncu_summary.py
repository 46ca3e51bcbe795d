#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: one block per launch with the metrics the roofline argument needs."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'sm__cycles_elapsed.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    for r in rows[2:]:
        print('---')
        for w, i in idx:
            print(f"  {w:66s} {r[i]:>22s} {units[i]}")


if __name__ == '__main__':
    main(sys.argv[1])
