#!/usr/bin/env python
"""Print the SASS of one address range of one kernel: sass_dump_loop.py obj name-substring lo hi"""
import re, sys, subprocess
obj, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
for f in re.split(r'\n\s+Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if pat not in name: continue
    for l in f.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m and lo <= int(m.group(1), 16) <= hi:
            print(f"{m.group(1)}  {m.group(2).strip()}")
