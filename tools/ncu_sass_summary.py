#!/usr/bin/env python
"""Summarises `ncu -i X.ncu-rep --page source --csv --print-source sass` output: executed instructions and stall samples
per opcode and the hottest instructions.  usage: ncu_sass_summary.py file.csv [n_top]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; data = rows[2:]
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); isamp = hdr.index('# Samples')
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[isamp]) for r in data)
print(rows[0][1]); print('total inst', tot, 'samples', tots)
op = collections.Counter(); ops = collections.Counter()
for r in data:
    s = r[isrc].strip().split()
    o = s[1] if s[0].startswith('@') else s[0]
    op[o.split('.')[0]] += int(r[ia]); ops[o.split('.')[0]] += int(r[isamp])
for o, c in op.most_common(25):
    print(f'{o:12s} {c:12d} {100*c/tot:5.1f}%  samples {100*ops[o]/tots:5.1f}%')
print()
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:ntop]
for i in sorted(top):
    print(i, data[i][isrc].strip()[:100], data[i][ia], data[i][isamp])
