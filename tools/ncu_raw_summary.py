#!/usr/bin/env python3
"""Print the handful of `ncu --page raw --csv` metrics that decide a bandwidth-bound kernel.

    ncu -i X.ncu-rep --page raw --csv | python tools/ncu_raw_summary.py
"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__cycles_elapsed.max",
    "launch__grid_size", "launch__block_size",
]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("---- " + r[idx["Kernel Name"]][:100])
    for w in WANT:
        if w in idx:
            print("  %-62s %18s %s" % (w, r[idx[w]], units[idx[w]]))
