#!/usr/bin/env python
"""Per-frame view of an ncu launch list of bench.py (--metrics gpu__time_duration.sum --csv): frames are split at
k_mesh_primary; prints each frame's traversal grid, per-kernel times and the traversal share, then the share over
the frames of bench.py's one-frame-at-a-time pass (traversal grids of 6 CTAs per SM).
usage: tools/launch_frames.py profiles/r1_launches_v7.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size")
seq = [(r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[-30:], float(r[vi].replace(",", "")) / 1000, r[gi])
       for r in rows[hdr + 1:] if len(r) > vi]
frames, cur = [], []
for k, v, g in seq:
    if "k_mesh_primary" in k:
        if cur:
            frames.append(cur)
        cur = []
    cur.append((k, v, g))
frames.append(cur)
names = ("k_mesh_primary", "k_shade<1>", "k_trace", "k_shade<0>", "k_tonemap")
print("frame  trace grid   primary  shade<1>   trace(2)  shade<0>(2)  tonemap   sum(us)  trace share")
full = []
for i, f in enumerate(frames[1:], 1):
    t = {n: sum(v for k, v, g in f if n in k) for n in names}
    grid = [g for k, v, g in f if "k_trace" in k]
    tot = sum(t.values())
    if grid and grid[0].startswith("(888"):
        full.append((t["k_trace"], tot))
    print("%5d  %-11s %8.1f %9.1f %10.1f %12.1f %8.1f %9.1f   %.3f" % (i, grid[0] if grid else "-", t["k_mesh_primary"], t["k_shade<1>"],
                                                                       t["k_trace"], t["k_shade<0>"], t["k_tonemap"], tot, t["k_trace"] / tot if tot else 0))
if full:
    tr, tot = sum(a for a, b in full[-2:]), sum(b for a, b in full[-2:])
    print("\ntimed steps (the last two frames with 888-CTA traversal grids): traversal %.1f us of %.1f us per frame = %.3f of the step"
          % (tr / 2, tot / 2, tr / tot))
