#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics we track from `ncu -i X.ncu-rep --page raw --csv` output."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'launch__registers_per_thread ',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum ', 'smsp__issue_active.avg.pct',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct', 'smsp__thread_inst_executed_per_inst',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum ',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'sass__inst_executed_shared_loads', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block ', 'launch__occupancy_limit', 'smsp__average_warps_issue_stalled', 'smsp__average_warp_latency_issue_stalled',
        'gpu__dram_throughput.avg.pct', 'lts__t_bytes.sum ', 'sm__cycles_elapsed.avg ', 'smsp__cycles_active.avg ']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("----", r[hdr.index("Kernel Name")][:100] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if any((h + ' ').startswith(k) or k.strip() in h and k.endswith('_') for k in KEYS) or any(h.startswith(k.strip()) for k in KEYS):
            print(f"{h:95s} {u:16s} {v}")
