#!/usr/bin/env python3
"""Instruction mix + top stall lines from an `ncu --page source --csv` export.  usage: ncu_mix.py file.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops, samp, tot, tots = collections.Counter(), collections.Counter(), 0, 0
lines = []
for r in rows[hi + 1:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    src, n, s = r[iS].strip(), int(r[iE]), int(r[iSamp] or 0)
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else src[:10]
    ops[op] += n
    samp[op] += s
    tot += n
    tots += s
    lines.append((s, n, src))
print('total warp instructions', tot, 'samples', tots)
for op, n in ops.most_common(22):
    print('%-10s %12d %5.1f%%   samples %6d %5.1f%%' % (op, n, 100.0 * n / tot, samp[op], 100.0 * samp[op] / max(tots, 1)))
print('--- top sampled instructions')
for s, n, src in sorted(lines, reverse=True)[:14]:
    print('%6d %10d  %s' % (s, n, src[:100]))
