#!/usr/bin/env python
"""Aggregates an `ncu --page source --csv` dump: stall reasons, opcode mix, hottest SASS ranges.
usage: srcstat.py src.csv [top_n]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; n = len(h)
data = [r for r in rows[2:] if len(r) >= n and r[0].startswith("0x")]
idx = {name: i for i, name in enumerate(h)}
st = [s for s in h if s.startswith('stall_') and 'Not Issued' not in s]
I = lambda r, k: int(float(r[idx[k]] or 0))
tot = collections.Counter()
for r in data:
    for s in st:
        tot[s] += I(r, s)
T = sum(tot.values()) or 1
print('stall reasons (all samples):')
for s, v in tot.most_common(10):
    print('  %-26s %8d %5.1f%%' % (s, v, 100 * v / T))
op = collections.Counter(); ops = collections.Counter()
for r in data:
    o = [x for x in r[idx['Source']].split() if not x.startswith('@')]
    o = o[0].split('.')[0] if o else '?'
    op[o] += I(r, 'Instructions Executed'); ops[o] += I(r, '# Samples')
TI = sum(op.values()) or 1; TS = sum(ops.values()) or 1
print('warp instructions %d, samples %d' % (TI, TS))
for o, v in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print('  %-10s inst %5.1f%%  samples %5.1f%%' % (o, 100 * v / TI, 100 * ops[o] / TS))
