#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: executed-instruction
histogram by region, top stall locations, stall-reason totals."""
import csv, sys, collections
path = sys.argv[1]
kernel_idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
# split into kernels
kernels = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; kernels.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    cur["rows"].append(r)
k = kernels[kernel_idx]
h = k["hdr"]; ix = {n: i for i, n in enumerate(h)}
print(k["name"], len(k["rows"]), "instructions")
base = int(k["rows"][0][0], 16)
tot_exec = sum(int(r[ix["Instructions Executed"]]) for r in k["rows"])
tot_samp = sum(int(r[ix["# Samples"]]) for r in k["rows"])
print("total warp-instructions executed", tot_exec, "samples", tot_samp)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.Counter()
for r in k["rows"]:
    for s in stalls:
        agg[s] += int(r[ix[s]])
print("stall samples:", [(s, v, round(100.0 * v / max(1, tot_samp), 1)) for s, v in agg.most_common(10)])
print("--- top 40 instructions by samples")
top = sorted(k["rows"], key=lambda r: -int(r[ix["# Samples"]]))[:40]
for r in top:
    off = int(r[0], 16) - base
    st = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:2]
    print(f"{off:05x} {r[1].strip():60s} samp={r[ix['# Samples']]:>6} exec={r[ix['Instructions Executed']]:>8} {st}")
if len(sys.argv) > 3:
    print("--- executed counts by instruction (offset, exec, text)")
    for r in k["rows"]:
        off = int(r[0], 16) - base
        print(f"{off:05x} {int(r[ix['Instructions Executed']]):>9} {int(r[ix['# Samples']]):>6} {r[1].strip()}")
