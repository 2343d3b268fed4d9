#!/usr/bin/env python
"""P_rand: the random-access roofline of this B200, measured as a sweep.

Independent random reads of aligned granules of 16 / 32 / 64 / 128 bytes over a table >> L2 (default 32 GiB), with
1 / 2 / 4 / 8 granules in flight per thread and 1..8 resident CTAs of 256 threads per SM
(mazu_b200_debug_gather_probe, include/mazu_b200_debug.h).  The maximum over the sweep per granule size is the peak;
DRAM bytes per granule (the line the memory system really moves) come from an ncu capture of the same launch
(profiles/r02_prand_ncu.txt).  Writes gpurun_out/r02_prand.json; the committed copy is profiles/r02_prand.json.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mazu_b200 as mz

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 32.0
quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
table = int(gib * (1 << 30))
n_items = 1 << (26 if quick else 28)
out = {"table_gib": gib, "n_items": n_items, "points": [], "best": {}}
for gran in (16, 32, 64, 128):
    best = None
    for ilp in (1, 2, 4, 8):
        for bps in ((8,) if quick else (1, 2, 4, 6, 8)):
            r = mz.gather_probe(table, n_items, granule_bytes=gran, ilp=ilp, blocks_per_sm=bps, iters=2)
            pt = {"granule_bytes": gran, "ilp": ilp, "blocks_per_sm": bps, "granules_per_s": r, "useful_gbs": r * gran / 1e9}
            out["points"].append(pt)
            if best is None or r > best["granules_per_s"]:
                best = pt
            print("gran %3d ilp %d ctas/SM %d: %.4g granules/s = %.1f GB/s useful" % (gran, ilp, bps, r, r * gran / 1e9), flush=True)
    out["best"][str(gran)] = best
mz.gather_probe(0, 0)  # free the table
# the line rate: a random granule of <= 128 B costs one 128-byte DRAM line (ncu), so lines/s = granules/s of the best point
out["lines_per_s"] = max(out["best"][g]["granules_per_s"] for g in out["best"])
out["sectors_per_s_32B"] = out["best"]["32"]["granules_per_s"]
out["line_gbs"] = out["lines_per_s"] * 128 / 1e9
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_prand.json"), "w"), indent=1)
print(json.dumps({k: out[k] for k in ("lines_per_s", "sectors_per_s_32B", "line_gbs")}))
