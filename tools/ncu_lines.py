#!/usr/bin/env python
"""Per-source-line instruction breakdown of an ncu report (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py report.ncu-rep [kernel-substring] [top-n]"""
import collections
import csv
import subprocess
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    i, seen = 0, set()
    agg = collections.OrderedDict()
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "File Path":
            path, fn, hdr = r[1], rows[i + 1][1], rows[i + 2]
            ci = {}
            for k, h in enumerate(hdr):
                ci.setdefault(h, k)
            j, cur = i + 3, None
            while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
                b = rows[j]
                j += 1
                if len(b) <= ci["Thread Instructions Executed"]:
                    continue
                if b[ci["Line No"]]:
                    cur = int(b[ci["Line No"]])
                if not b[ci["Address"]]:
                    continue
                a = agg.setdefault(fn, collections.defaultdict(lambda: [0.0, 0.0, 0.0]))[(path.split("/")[-1], cur)]
                a[0] += num(b[ci["Instructions Executed"]])
                a[1] += num(b[ci["Thread Instructions Executed"]])
                a[2] += num(b[ci["# Samples"]])
            i = j
        else:
            i += 1
    for fn, lines in agg.items():
        if pat not in fn or fn in seen:
            continue
        seen.add(fn)
        tot = sum(v[0] for v in lines.values())
        thr = sum(v[1] for v in lines.values())
        print(f"===== {fn[:110]}\n      warp inst {tot / 1e6:.1f} M, avg active threads {thr / max(tot, 1):.1f}")
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"{f:14s}:{ln or 0:4d}  {v[0] / tot * 100:6.2f}%  avg_thr {v[1] / max(v[0], 1):5.1f}  stall samples {v[2]:.0f}")


if __name__ == "__main__":
    main()
