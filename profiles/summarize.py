#!/usr/bin/env python3
"""Write a text summary of an ncu report: key raw metrics + stall reasons + hottest CUDA source lines.
usage: summarize.py report.ncu-rep > summary.txt"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max"]
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name[:140])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("  %-70s %-10s %s" % (k, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = [r for r in rows if "Instructions Executed" in r][0]
data = rows[rows.index(h) + 1:]
cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = collections.Counter()
for r in data:
    for i in cols:
        if i < len(r) and r[i] not in ("", "-"):
            tot[h[i]] += int(r[i])
s = sum(tot.values()) or 1
print("stall reasons (all samples):", ", ".join("%s %.1f%%" % (k, 100.0 * v / s) for k, v in tot.most_common(8)))
print("SASS instructions in kernel:", len(data))
print("hottest CUDA source lines (share of warp instructions / share of stall samples):")
sys.stdout.flush()
subprocess.run([sys.executable, __file__.replace("summarize.py", "ncu_lines.py"), rep, "2.0"])
