#!/usr/bin/env python
"""Summarise an ncu report (or launch-list csv) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/r01a/prof_vr.ncu-rep > profiles/r01_vr_northstar.txt
    python tools/ncu_summary.py --launches gpurun_out/r01a/launches.csv > profiles/r01_launches.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = """gpu__time_duration.sum
sm__cycles_elapsed.avg.per_second
launch__grid_size
launch__block_size
launch__registers_per_thread
launch__shared_mem_per_block
launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem
sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio
smsp__issue_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
dram__bytes_read.sum
dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct
lts__t_bytes.sum
lts__t_bytes.sum.per_second
lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__m_xbar2l1tex_read_bytes.sum
l1tex__m_xbar2l1tex_read_bytes.sum.per_second
dram__bytes_read.sum.per_second
l1tex__t_sector_hit_rate.pct
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
sass__inst_executed_shared_loads
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio""".split()


def report(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none, raw page of", path)
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS + [h for h in hdr if "data_pipe_lsu_wavefronts" in h and h not in KEYS]:
            if k in hdr:
                i = hdr.index(k)
                print("  %-92s %-16s %s" % (k, units[i], r[i]))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "second": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki][:110], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold cache: compare shares)")
    print("# source:", path, " total %.3f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%10.3f ms %5d launches %6.2f%%  avg %9.4f ms  %s" % (t, n, 100 * t / tot, t / n, k))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        for p in sys.argv[1:]:
            report(p)
