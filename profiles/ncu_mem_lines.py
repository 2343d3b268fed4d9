#!/usr/bin/env python3
"""Per CUDA source line of an ncu report: executed global-memory instructions, L2 sectors requested, share of long-scoreboard stalls.
usage: ncu_mem_lines.py report.ncu-rep [units] -- `units` (e.g. lookups in the launch) scales the counts per unit"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        def g(name):
            if name not in hdr: return 0
            v = r[hdr.index(name)].split("(")[0].strip()
            try: return int(v)
            except ValueError: return 0
        key = (cur, int(r[0]))
        a = agg[key]
        a[0] += g("Instructions Executed"); a[1] += g("L2 Theoretical Sectors Global"); a[2] += g("stall_long_sb"); a[3] += g("L1 Tag Requests Global"); a[4] = r[1].strip()[:90]
tot_sec = sum(a[1] for a in agg.values()) or 1
tot_lsb = sum(a[2] for a in agg.values()) or 1
print("total L2 theoretical sectors (global) %d%s; long_sb samples %d" % (tot_sec, (" = %.2f per unit" % (tot_sec / units)) if units else "", tot_lsb))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[1] * 100.0 / tot_sec >= 1.0 or a[2] * 100.0 / tot_lsb >= 2.0:
        per = (" %.3f sec/unit" % (a[1] / units)) if units else ""
        print("%5.1f%% sectors%s %5.1f%% long_sb  %s:%d: %s" % (100.0 * a[1] / tot_sec, per, 100.0 * a[2] / tot_lsb, f, ln, a[4]))
