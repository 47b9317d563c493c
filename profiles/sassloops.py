#!/usr/bin/env python
"""Static look at a cuobjdump -sass listing: loops (backward branches) with instruction / FP64 / spill / shared counts.
usage: cuobjdump -sass -fun <mangled> lib.so > k.sass; sassloops.py k.sass"""
import re, sys
ins = []
for l in open(sys.argv[1]):
    m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
print(len(ins), 'instructions')
loops = []
for a, t in ins:
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
op = lambda t, names: re.match(r'(@\S+\s+)?(' + names + r')\b', t) is not None
def cnt(lo, hi):
    sel = [t for a, t in ins if lo <= a <= hi]
    return dict(n=len(sel), fp64=sum(op(t, 'DFMA|DMUL|DADD') for t in sel), LDL=sum(op(t, r'LDL\S*') for t in sel),
                STL=sum(op(t, r'STL\S*') for t in sel), LDS=sum(op(t, r'LDS\S*') for t in sel), STS=sum(op(t, r'STS\S*') for t in sel),
                LDG=sum(op(t, r'LDG\S*') for t in sel), STG=sum(op(t, r'STG\S*') for t in sel))
for lo, hi in sorted(loops):
    print('%#7x-%#7x' % (lo, hi), cnt(lo, hi))
print('total', cnt(0, 1 << 30))
