#!/usr/bin/env python3
"""Measured pinned-memory copy bandwidth of the box (the bound of bench.py's e2e figure): D2H alone, H2D alone and both at once.
usage: pcie_probe.py [GiB]   -> one JSON line"""
import json, sys
import torch
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * (1 << 30))
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        torch.cuda.synchronize()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 1e3)
    return best
def d2h():
    h.copy_(d, non_blocking=True)
def h2d():
    d2.copy_(h2, non_blocking=True)
def both():
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
for f in (d2h, h2d, both):
    f()
t1, t2, t3 = timed(d2h), timed(h2d), timed(both)
print(json.dumps({"gib": gib, "d2h_gbs": n / t1 / 1e9, "h2d_gbs": n / t2 / 1e9, "duplex_each_gbs": n / t3 / 1e9}))
