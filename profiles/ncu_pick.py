#!/usr/bin/env python
"""Print selected metrics of an `ncu --page raw --csv` export (helper for profiles/ summaries)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = sys.argv[2:] or ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64', 'smsp__issue_active.avg.pct',
    'launch__occupancy_limit', 'smsp__average_warp', 'smsp__average_warps_issue_stalled',
    'l1tex__t_sector_hit_rate', 'lts__t_sector_hit_rate', 'smsp__thread_inst_executed_per_inst',
    'launch__waves', 'sm__throughput.avg.pct', 'local', 'launch__grid_size', 'launch__block_size',
    'sm__pipe_fp64', 'smsp__inst_executed_pipe_fp64', 'gpu__dram_throughput', 'shared_mem']
for vals in rows[2:]:
    print('---', vals[hdr.index('Kernel Name')][:70] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, units, vals):
        if any(w in h for w in want):
            print(f"{h:95s} {u:14s} {v}")
