#!/usr/bin/env python
"""ncu -i <rep> --page raw --csv  ->  a short per-kernel summary of the metrics the roofline uses.
Usage: python profiles/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x.summary.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"kernel: {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
    for h, u in zip(hdr, units):
        if h in WANT:
            print(f"  {h:75s} {d[h]:>18s} {u}")
