#!/usr/bin/env python
"""Measure P_rand: independent random 32-byte gathers over a table much larger than L2."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mazu_b200 as mz
out = {}
for gib in (1, 8, 32):
    r = mz.measure_random_gather(gib << 30, 1 << 30, iters=3)
    out["%dGiB" % gib] = {"sectors_per_s": r, "GBps": r * 32 / 1e9}
    print(gib, "GiB table: %.3g sectors/s = %.1f GB/s" % (r, r * 32 / 1e9), flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "prand.json"), "w"), indent=1)
