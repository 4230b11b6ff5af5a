#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total time, share.
usage: launch_list_summary.py <launches.csv> <out.md> "<command that was profiled>" """
import csv
import re
import sys
from collections import OrderedDict

src, dst, what = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for row in r:
    if len(row) <= vi:
        continue
    name = re.sub(r"\(.*", "", row[ki]).strip()
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list summary: {what}", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none` over the whole command. Per-launch times under ncu are cold-cache and "
       "serialised (no overlap between consecutive steps); the SHARE is what matters. Raw list: `" + src.split("/")[-1] + "`.", "",
       "| kernel | launches | total us | share |", "|---|---|---|---|"]
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {name} | {n} | {us:.1f} | {100 * us / tot:.1f} % |")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
