#!/usr/bin/env python
"""List the loops (backward branches) of each kernel in a cuobjdump -sass dump with their
instruction mix, so the hot loop's instruction count can be tracked without a GPU."""
import re, sys, collections, subprocess
obj = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
for f in re.split(r'\n\s+Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if pat and pat not in name: continue
    ins = []
    for l in f.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name, len(ins), "instructions")
    for a, t in ins:
        m = re.search(r'\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            b = int(m.group(1), 16)
            body = [(x, u) for x, u in ins if b <= x <= a]
            c = collections.Counter()
            for x, u in body:
                u = re.sub(r'^@!?U?P\d\s+', '', u)
                c[u.split()[0].split('.')[0]] += 1
            cold = sum(1 for x, u in body if 'CALL' in u)
            print(f"  loop {b:#x}..{a:#x}: {len(body)} instrs, calls={cold}, local={c.get('STL',0)+c.get('LDL',0)}")
            print("    ", c.most_common(24))

def hot_path(ins, lo, hi):
    """Instructions on the path through loop [lo,hi] that skips every forward-branch region
    containing a CALL (the cold fallbacks) and falls through all other conditional branches."""
    addr = {a: k for k, (a, t) in enumerate(ins)}
    k = addr[lo]; n = 0; mix = collections.Counter()
    while True:
        a, t = ins[k]
        n += 1
        op = re.sub(r'^@!?U?P\d\s+', '', t).split()[0].split('.')[0]
        mix[op] += 1
        if a == hi: break
        m = re.search(r'\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            cond = t.startswith('@') or re.search(r'BRA(?:\.U)?\s+!?U?P\d', t)
            if tgt > a:
                region = [u for x, u in ins if a < x < tgt]
                if not cond or any('CALL' in u for u in region):
                    k = addr[tgt]; continue
        k += 1
    return n, mix

if len(sys.argv) > 3:
    for f in re.split(r'\n\s+Function : ', txt)[1:]:
        name = f.split('\n')[0]
        if pat and pat not in name: continue
        ins = []
        for l in f.split('\n'):
            m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
            if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
        n, mix = hot_path(ins, lo, hi)
        print("hot path", n, mix.most_common(30))
