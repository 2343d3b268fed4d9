#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line: share of executed warp instructions and of stall samples.
usage: ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "Line No"):
        ia = hdr.index("Instructions Executed"); ismp = hdr.index("# Samples"); ith = hdr.index("Avg. Threads Executed")
        try: agg.append((cur, int(r[0]), r[1], int(r[ia] or 0), int(r[ismp] or 0), r[ith]))
        except ValueError: pass
tot = sum(a[3] for a in agg); tots = sum(a[4] for a in agg)
print("total warp instructions %d, stall samples %d" % (tot, tots))
for f, ln, src, n, k, th in sorted(agg, key=lambda a: -a[4]):
    if 100.0 * n / tot >= minpct or 100.0 * k / tots >= minpct:
        print("%5.1f%% inst %5.1f%% smp thr/inst %-4s %s:%d: %s" % (100.0 * n / tot, 100.0 * k / tots, th, f, ln, src.strip()[:100]))
