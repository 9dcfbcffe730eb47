#!/usr/bin/env python
"""All `pct_of_peak` metrics of an ncu report's raw page, sorted, plus the stall breakdown: a quick look at what is full."""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")])
    items = []
    for h, u, v in zip(hdr, units, r):
        if "pct_of_peak" in h or h.endswith(".pct"):
            try:
                items.append((float(v.replace(",", "")), h))
            except ValueError:
                pass
    for v, h in sorted(items, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
        print("  %8.2f  %s" % (v, h))
    for h, u, v in zip(hdr, units, r):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") or h in (
                "gpu__time_duration.sum", "dram__bytes_read.sum", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum",
                "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "launch__registers_per_thread"):
            print("  %-90s %s %s" % (h, v, u))
