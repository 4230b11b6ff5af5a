#!/usr/bin/env python
"""Top SASS instructions of an ncu report by samples of one stall reason (default: all samples), with neighbours.
usage: ncu_hot.py report.ncu-rep [stall_column_substring] [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; key = sys.argv[2] if len(sys.argv) > 2 else "# Samples"; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Kernel Name": hdr = None; data = []; continue
    if hdr is None and r and "Source" in r: hdr = r; continue
    if hdr: data.append(r)
ix = {n: i for i, n in enumerate(hdr)}
col = [n for n in hdr if key in n][0]
tot = sum(int(r[ix[col]] or 0) for r in data)
print("column:", col, "total", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda k: -int(data[k][ix[col]] or 0))[:topn]
for k in order:
    r = data[k]
    print(f"{int(r[ix[col]]):6d}  #{k:5d} exec={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:110]}")
