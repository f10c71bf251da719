#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line.
usage: ncu_lines.py export.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, agg = None, {}
def num(s):
    try: return int(s)
    except Exception: return 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if not r or r[0] in ('Line No', 'Function Name'): continue
    if r[0].isdigit() and len(r) >= 8 and r[2] == '-':
        a = agg.setdefault((cur, int(r[0])), [0, 0, r[1]])
        a[0] += num(r[4]); a[1] += num(r[7])
tot = sum(a[0] for a in agg.values()) or 1
print("total samples", tot, " total warp-inst", sum(a[1] for a in agg.values()))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:5d} {a[0]:6d} {100*a[0]/tot:5.1f}% inst={a[1]:8d} | {a[2].strip()[:100]}")
