#!/usr/bin/env python3
"""Summarise an ncu report for profiles/: selected raw metrics per captured launch.
usage: python tools/ncu_summary.py gpurun_out/r1_stream.ncu-rep > profiles/r1_price_lv_stream_ncu_summary.txt"""
import csv, io, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed_op_global_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
kcol = names.index("Kernel Name")
print("# %s: %d launch(es): %s" % (rep, len(data), " | ".join(sorted(set(r[kcol].split("(")[0] for r in data)))))
for m in WANT:
    if m in names:
        c = names.index(m)
        print("%-70s %-12s %s" % (m, units[c], " | ".join(r[c] for r in data)))
