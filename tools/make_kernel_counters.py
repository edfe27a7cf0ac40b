#!/usr/bin/env python
"""profiles/<name>.txt (tools/ncu_summary.py output) -> profiles/kernel_counters.json, the ncu-measured counters bench.py
attaches to its roofline blocks (roofline.traffic / dram_frac / issue_frac / lanes_per_warp / l1,l2 hit rates).
usage: tools/make_kernel_counters.py workload=profiles/file.txt[:kernel substring] ..."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def parse(path, want):
    launches, cur = [], None
    for line in open(path):
        if line.startswith("## "):
            cur = {"kernel": line[3:].strip()}
            launches.append(cur)
        elif cur is not None and line.startswith("  "):
            m = re.match(r"\s+(.*?)\s{2,}([-\d.,]+)\s*(\S*)\s*$", line)
            if m:
                cur[m.group(1).strip()] = (float(m.group(2).replace(",", "")), m.group(3))
    sel = [l for l in launches if want in l["kernel"]] if want else launches
    if not sel:
        raise SystemExit(f"{path}: no launch of '{want}'")

    def mean(key, scale=False):
        vals = []
        for l in sel:
            if key in l:
                v, u = l[key]
                vals.append(v * UNIT.get(u, 1.0) if scale else v)
        return sum(vals) / len(vals) if vals else None
    return {"kernel": sel[0]["kernel"][:80], "launches_captured": len(sel),
            "dram_bytes_per_launch": (mean("DRAM read", True) or 0) + (mean("DRAM write", True) or 0),
            "l1_hit_rate": mean("L1 hit rate %"), "l2_hit_rate": mean("L2 hit rate %"),
            "lanes_per_warp": mean("warp execution efficiency (active threads / 32-wide instr)"),
            "issue_slot_utilisation": (mean("issue-slot utilisation %") or 0) / 100.0,
            "achieved_occupancy_pct": mean("achieved occupancy %"), "warp_instructions": mean("warp instructions"),
            "duration_under_ncu_us": mean("duration", False),
            "source": os.path.relpath(path, ROOT) + " (ncu --set full --clock-control none; cold caches, serialised replays)"}


def main():
    out_path = os.path.join(ROOT, "profiles", "kernel_counters.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for arg in sys.argv[1:]:
        wl, rest = arg.split("=", 1)
        path, _, want = rest.partition(":")
        out[wl] = parse(path, want)
    json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
