#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: share / count / average per kernel, and per
grid for the kernels matching --grid.  Usage: python scripts/launch_shares.py gpurun_out/launches.csv [--grid conv_tc]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    pat = sys.argv[sys.argv.index("--grid") + 1] if "--grid" in sys.argv else None
    rows = list(csv.reader(open(path, errors="ignore")))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.defaultdict(lambda: [0, 0.0])
    grids = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 2:]:
        if len(r) <= vi:
            continue
        try:
            t = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("aotb::", "")
        agg[n][0] += 1
        agg[n][1] += t
        if pat and pat in n:
            grids[(n, r[gi])][0] += 1
            grids[(n, r[gi])][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"total kernel time {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:18]:
        print(f"{t / tot * 100:5.1f}%  n={c:5d}  avg={t / c / 1000:8.1f} us  {n[:80]}")
    for k, (c, t) in sorted(grids.items(), key=lambda x: -x[1][1]):
        print(f"   {k[0][:40]:40s} grid {k[1]:14s} n={c:4d} avg={t / c / 1000:7.1f} us  total={t / 1e3:8.1f} us")


if __name__ == "__main__":
    main()
